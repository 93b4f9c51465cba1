"""extract_index host logic (reference extract_index.py:30-58) on CPU: the clip visit order equals the real
shuffled DataLoader's under the same seed, the stop rule and the stride subsampling agree with the reference loop."""
import torch

from tinyvc_b200 import index as ix


class _DS(torch.utils.data.Dataset):
    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return torch.tensor(i), torch.tensor(0.0)


def test_visit_order_equals_shuffled_dataloader():
    for n, seed in ((1, 0), (7, 1), (100, 2), (1000, 3)):
        torch.manual_seed(seed)
        want = [int(i) for i, _ in torch.utils.data.DataLoader(_DS(n), batch_size=1, shuffle=True)]
        after_ref = torch.rand(1)
        torch.manual_seed(seed)
        got = ix.dataloader_order(n)
        after_ours = torch.rand(1)
        assert got == want
        assert torch.equal(after_ref, after_ours), "global RNG left in a different state than the reference loop leaves it"


def test_stop_rule_and_stride():
    assert ix.columns_per_clip(48000, 4) == 25 and ix.columns_per_clip(48000, 1) == 100 and ix.columns_per_clip(4800, 4) == 3
    lengths = [48000] * 200
    order = list(range(199, -1, -1))
    picked = ix.clips_needed(lengths, order, size=2048, stride=4)
    assert picked == order[:82]            # 81 clips give 2025 <= 2048; the 82nd pushes the total past size
    assert ix.clips_needed(lengths[:3], [2, 0, 1], size=2048, stride=4) == [2, 0, 1]   # small dataset: everything


def test_cli_flags_match_reference():
    import extract_index
    a = vars(extract_index.build_parser().parse_args([]))
    assert a == dict(dataset_cache="dataset_cache", encoder_path="models/encoder.pt", size=2048, output="models/index.pt",
                     device="cuda", stride=4)
