"""Per-layer cost model of the FilterNet convs at a given (B, Lf), next to the measured per-kernel times of a bench line.

For every tc_conv launch of the Down / Up blocks it lists the tiling tc_conv_launch picks (mode, tiles, K-stages, smem ring),
the bytes a CTA pulls through L2 (activation windows + weight images, weights-resident mode included), the MMA count,
and three lower bounds in microseconds:
    mma   = MMAs per CTA x 57 cycles (measured issue cost per tcgen05.mma for N <= 64; N = 128 counted double)
    load  = bytes per CTA / 39 B/cycle (TMA tensor copy) or / 12 B/cycle (per-thread cp.async: edge and flat tiles)
    dram  = unique bytes of the layer (in + out + weights) / 6.5 TB/s
so that `measured / max(bounds)` says how far a layer is from what its own tiling allows, and `load / dram` how much
of its traffic is re-fetching (weights per row tile) -- the levers of the next round.

    python tools/layer_model.py [bench.json]          # default profiles/r01z_bench_shipped_default.json
"""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMS, GHZ = 148, 1.9
UP_CH, UP_OUT, UP_FAC = [384, 192, 96, 48, 24], [192, 96, 48, 24, 24], [2, 3, 4, 4, 5]
DN_IN, DN_OUT, DN_FAC = [24, 48, 96, 192], [48, 96, 192, 384], [5, 4, 4, 3]
DN_NT12, DN_NT3 = [24, 48, 32, 32], [48, 48, 64, 48]
UP_NT, UP_NT5 = [48, 64, 48, 48, 24], [32, 32, 48, 24, 24]
SMEM = 200 * 1024


def align(v, a):
    return (v + a - 1) // a * a


def pick_kb(cin):
    c16 = align(cin, 16)
    if c16 <= 64:
        return c16
    if c16 % 64 == 0:
        return 64
    if c16 % 48 == 0:
        return 48
    return 64


def layer(name, B, T, cin, cout, taps, dil, nt, aux_cin=0, film=False, out_bytes_per_ch=4, extra_in=0):
    kb = pick_kb(max(cin, aux_cin))
    nkb = math.ceil(align(cin, 16) / kb)
    aux_nkb = math.ceil(align(aux_cin, 16) / kb) if aux_cin else 0
    ntp = align(nt, 16)
    n_tiles = math.ceil(cout / nt)
    tpu = math.ceil(T / 128)
    halo = taps == 3 and T / (tpu * 128) >= 0.75
    row_tiles = B * tpu if halo else math.ceil(B * T / 128)
    tiles = row_tiles * n_tiles
    grid = min(tiles, SMS)
    per_cta = math.ceil(tiles / grid)
    R = 128 + 2 * dil if halo else 128
    g_main = (R + 14) // 8 if halo else 16
    g_aux = 17 if halo else 16
    lbo = max(g_main, g_aux if aux_cin else 0) * 128
    a_stage = align(2 * (kb // 8) * lbo, 128)
    b_main = 4 * kb * ntp * (taps if halo else 1)
    b_aux = 4 * kb * (2 * ntp if film else ntp) if aux_cin else 0
    image = (4 * kb * ntp * taps * nkb + b_aux * aux_nkb)
    wres = n_tiles == 1 and image <= 64 * 1024
    stage = a_stage + (0 if wres else max(b_main, b_aux))
    ring = min(8, (SMEM - 256 - (image if wres else 0)) // stage)
    n_main = nkb if halo else taps * nkb
    stages = n_main + aux_nkb
    ks = kb // 16
    mma = 3 * ks * (taps * nkb + aux_nkb)
    nfac = 2.0 if ntp > 64 else 1.0
    # bytes per tile through L2 -> SM
    a_bytes = n_main * 2 * (kb // 8) * (g_main if halo else 16) * 128 + aux_nkb * 2 * (kb // 8) * g_aux * 128
    w_bytes = 0 if wres else image
    edge_frac = 1.0 if not halo else min(1.0, 2.0 / tpu)         # tiles touching an utterance edge (cp.async gathers)
    tma = taps == 1 or halo
    bpc = 39.0 * (1 - edge_frac) + 12.0 * edge_frac if tma else 12.0
    t_mma = per_cta * mma * 57 * nfac / (GHZ * 1e3)
    t_load = (per_cta * (a_bytes + w_bytes) + (image if wres else 0)) / bpc / (GHZ * 1e3)
    rows = B * T
    uniq = rows * align(cin, 8) * 4 + rows * align(aux_cin, 8) * 4 + rows * cout * out_bytes_per_ch + image * n_tiles + extra_in * rows
    t_dram = uniq / 6.5e6
    l2_total = tiles * (a_bytes + w_bytes)
    return dict(name=name, mode="halo" if halo else "flat", tiles=tiles, per_cta=per_cta, stages=stages, ring=ring, wres=wres,
                mma_per_tile=mma, kb_per_tile=(a_bytes + w_bytes) / 1024, l2_mb=l2_total / 1e6, uniq_mb=uniq / 1e6,
                t_mma=t_mma, t_load=t_load, t_dram=t_dram)


def main() -> int:
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r01z_bench_shipped_default.json")
    d = json.load(open(path))
    meas = d["roofline"]["per_kernel_ms_per_step"]
    B, Lf = 64, 18
    L = Lf * 480
    rows = []
    rows.append(layer("tc_down0", B, L, 17, 24, 3, 1, 24, out_bytes_per_ch=8))
    T = L
    for i in range(4):
        T //= DN_FAC[i]
        rows.append(layer(f"tc_down{i+1}_c1", B, T, DN_IN[i], DN_IN[i], 3, 1, DN_NT12[i]))
        rows.append(layer(f"tc_down{i+1}_c2", B, T, DN_IN[i], DN_IN[i], 3, 2, DN_NT12[i]))
        rows.append(layer(f"tc_down{i+1}_c3", B, T, DN_IN[i], DN_OUT[i], 3, 4, DN_NT3[i], aux_cin=DN_IN[i], out_bytes_per_ch=8))
    T = Lf
    for i in range(5):
        T *= UP_FAC[i]
        c = UP_CH[i]
        rows.append(layer(f"tc_up{i}_c1", B, T, c, c, 3, 1, UP_NT[i]))
        rows.append(layer(f"tc_up{i}_c2", B, T, c, c, 3, 3, UP_NT[i], aux_cin=c, film=True, out_bytes_per_ch=8, extra_in=4 * c))
        rows.append(layer(f"tc_up{i}_c3", B, T, c, c, 3, 9, UP_NT[i]))
        rows.append(layer(f"tc_up{i}_c4", B, T, c, c, 3, 27, UP_NT[i], aux_cin=c, film=True, extra_in=4 * c))
        rows.append(layer(f"tc_up{i}_c5", B, T, c, UP_OUT[i], 1, 1, UP_NT5[i]))
    print(f"{'layer':14s} {'mode':4s} {'tiles':>6s} {'/cta':>4s} {'stg':>3s} {'ring':>4s} {'wres':>4s} {'mma/t':>5s} {'KB/t':>6s} {'L2 MB':>7s} {'uniq MB':>7s} "
          f"{'mma us':>6s} {'load us':>7s} {'dram us':>7s} {'meas us':>7s} {'meas/bound':>10s}")
    tot_m = tot_b = 0.0
    for r in rows:
        m = meas.get(r["name"], float("nan")) * 1e3
        bound = max(r["t_mma"], r["t_load"], r["t_dram"])
        tot_m += m
        tot_b += bound
        print(f"{r['name']:14s} {r['mode']:4s} {r['tiles']:6d} {r['per_cta']:4d} {r['stages']:3d} {r['ring']:4d} {str(r['wres'])[0]:>4s} {r['mma_per_tile']:5d} "
              f"{r['kb_per_tile']:6.0f} {r['l2_mb']:7.1f} {r['uniq_mb']:7.1f} {r['t_mma']:6.1f} {r['t_load']:7.1f} {r['t_dram']:7.1f} {m:7.1f} {m / bound:10.1f}")
    print(f"sum measured {tot_m:.0f} us, sum of per-layer bounds {tot_b:.0f} us  (bench step {d['ms_per_step'] * 1e3:.0f} us: profiling pass adds event overhead)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
