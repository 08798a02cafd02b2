"""Host-side launch planning of the tcgen05 convolution kernels, checked without a GPU: every convolution of the
hybrid net (hybridnet.py:379-423) at the headline 512x512x48 shape, in every pass (fprop / dgrad / wgrad) and both
tensor-core precisions (1 = bf16, 2 = bf16x3), must get a plan that respects the SM's limits -- 227 KB of dynamic
shared memory, 512 TMEM columns, UMMA N in [16, 256] and a multiple of 16 -- and whose packed-weight workspace is
what hdn_conv_tc_workspace reports.  hdn_conv_tc_plan is host arithmetic only (include/hdn.h)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))

SMEM_MAX = 227 * 1024


@pytest.fixture(scope="module")
def plan_rows():
    import h_denseunet_b200._lib as L
    L.build()
    import plan_table
    return {prec: plan_table.plans(prec) for prec in (1, 2)}


@pytest.mark.parametrize("prec", [1, 2])
def test_every_convolution_has_a_valid_plan(plan_rows, prec):
    rows = plan_rows[prec]
    assert len(rows) == 3 * 231                                  # 231 convolutions x 3 passes
    fallbacks = [(n, ps) for n, ps, o, _, _ in rows if o is None]
    # only the two 3-class classifiers' dgrad / wgrad (Cout = 3) stay on the fp32 FMA kernels
    assert sorted(fallbacks) == sorted([("dense167classifer", 1), ("dense167classifer", 2),
                                        ("2d3dclassifer", 1), ("2d3dclassifer", 2)]), fallbacks
    for name, ps, o, ws, M in rows:
        if o is None:
            continue
        bn, tiles, kb, ck, ring, nraw, tmem, smem, flat, P, x3, s2d, work, fit, k_or_bne, nc_or_ci = o
        what = (name, ps, o)
        assert 0 < smem <= SMEM_MAX, what
        assert tmem in (32, 64, 128, 256, 512), what
        assert bn % 16 == 0 and 16 <= bn <= 256, what
        assert work > 0 and fit == 1, what
        assert x3 == (1 if prec == 2 else 0), what
        if ps < 2:
            assert 2 * bn <= tmem, what                            # two accumulator buffers
            assert ring >= 2, what                                 # weight ring
            assert s2d or nraw >= 2, what                          # raw fp32 ring (SIMT producers) / A-stage ring (TMA mode)
            assert ck == (32 if prec == 2 else ck) and ck in (32, 64), what
            nsplit = 2 if prec == 2 else 1
            wbytes = tiles * kb * bn * ck * 2 * nsplit
            assert ws > 0 and (ws % (bn * ck * 2 * nsplit) == 0 or ws > wbytes), what
            if P in (180, 128) and not s2d:
                # TMA mode (default for every stride-1 layer): packed weights + the pre-packed bf16 head (+ tail) operand
                k = k_or_bne
                opb = (M * k * 2 + 255) // 256 * 256
                assert ws >= opb * nsplit and (ws - opb * nsplit) % 256 == 0, what
        else:
            bne, ci = k_or_bne, nc_or_ci
            assert bne == bn * (2 if prec == 2 else 1) and bne <= 256, what
            assert ci == (64 if prec == 2 else 128), what
            assert ring * bne <= tmem, what                        # one accumulator per tap of the group
            if ws:
                # second-generation kernel (conv_tc2_wgrad.cu): plain bf16 only, 1x3x3 / 3x3x3; scratch = the two bf16
                # operand tensors of the pre-pass; `nraw` reports the depth of the TMA stage ring
                assert prec == 1 and P in (180, 160, 128) and 2 <= nraw <= 6, what
                assert ws % 256 == 0, what
            else:
                assert prec == 2 or s2d, what                      # gen 1 keeps bf16x3 and the stems


def test_bf16x3_workspace_is_twice_the_bf16_one_per_channel_block(plan_rows):
    """bf16x3 packs a head and a tail block per (column tile, K block, tap): same bytes per covered channel x2."""
    # (layers in the SIMT-producer form: the stems; the TMA-mode layers add their operand tensors to the scratch)
    a = {(n, ps): (o, ws) for n, ps, o, ws, _ in plan_rows[1] if o is not None and ps < 2 and (o[11] or o[9] == 209)}
    b = {(n, ps): (o, ws) for n, ps, o, ws, _ in plan_rows[2] if o is not None and ps < 2 and (o[11] or o[9] == 209)}
    assert a.keys() == b.keys()
    for key in a:
        (o1, w1), (o2, w2) = a[key], b[key]
        cover1 = o1[0] * o1[1] * o1[2] * o1[3]                     # BN * column tiles * K blocks * CK
        cover2 = o2[0] * o2[1] * o2[2] * o2[3]
        assert w1 * cover2 * 2 == w2 * cover1, (key, o1, w1, o2, w2)
