from tinyvc_b200.tinyvc import Decoder, Encoder, match_features  # noqa: F401
from tinyvc_b200.tinyvc import decoder, encoder, convnext, feature_retrieval  # noqa: F401
