"""Golden vectors for llama3_flash_attn_prepare_cu_seqlens (reference
test/test_llama3_prepare_cu_seqlens.py:7-26)."""
import torch

from ring_flash_attn_b200 import llama3_flash_attn_prepare_cu_seqlens as prepare

GOLD_CAUSAL = {
    0: ([0, 2], [0, 2], 2, 2, (0, 2)),
    1: ([0, 2], [0, 4], 2, 4, (0, 4)),
    2: ([0, 2], [0, 6], 2, 6, (0, 6)),
    3: ([0, 1, 2], [0, 7, 8], 1, 7, (0, 8)),
    4: ([0, 2], [0, 3], 2, 3, (7, 10)),
    5: ([0, 2], [0, 5], 2, 5, (7, 12)),
    6: ([0, 2], [0, 7], 2, 7, (7, 14)),
    7: ([0, 2], [0, 2], 2, 2, (14, 16)),
}


def test_golden_causal():
    cu = torch.tensor([0, 7, 14, 16], dtype=torch.int32)
    for rank, (cq, ck, mq, mk, sl) in GOLD_CAUSAL.items():
        g_cq, g_ck, g_mq, g_mk, g_sl = prepare(cu, True, rank, 8)
        assert g_cq.tolist() == cq and g_ck.tolist() == ck
        assert (g_mq, g_mk) == (mq, mk)
        assert (g_sl.start, g_sl.stop) == sl
        assert g_mq == (g_cq[1:] - g_cq[:-1]).max().item()
        assert g_mk == (g_ck[1:] - g_ck[:-1]).max().item()


def test_golden_noncausal():
    cu = torch.tensor([0, 7, 14, 16], dtype=torch.int32)
    for rank, ck, sl in [(0, [0, 7], (0, 7)), (3, [0, 7, 14], (0, 14)), (6, [0, 7], (7, 14))]:
        _, g_ck, _, _, g_sl = prepare(cu, False, rank, 8)
        assert g_ck.tolist() == ck and (g_sl.start, g_sl.stop) == sl


def test_golden_long():
    cu = torch.tensor([0, 120, 1248, 4232], dtype=torch.int32)
    gold = {
        0: ([0, 120, 529], [0, 120, 529], (0, 529)),
        1: ([0, 529], [0, 938], (120, 1058)),
        2: ([0, 190, 529], [0, 1128, 1467], (120, 1587)),
        7: ([0, 529], [0, 2984], (1248, 4232)),
    }
    for rank, (cq, ck, sl) in gold.items():
        g_cq, g_ck, _, _, g_sl = prepare(cu, True, rank, 8)
        assert g_cq.tolist() == cq and g_ck.tolist() == ck and (g_sl.start, g_sl.stop) == sl
