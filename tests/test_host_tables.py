"""Integer / 0-1 host tables must be bit-exact with the reference (fixtures produced by the reference's own
seq_aligner / utils / attention_control functions, oracle/make_golden.py `tables`)."""
import json
import os

import pytest
import torch

from pnpinversion_b200 import ptp_utils, scheduler, seq_aligner, synth
from pnpinversion_b200.attention_control import get_equalizer

GOLD = os.path.join(os.path.dirname(__file__), "golden", "tables.json")


@pytest.fixture(scope="module")
def gold():
    with open(GOLD) as f:
        return json.load(f)


def test_refinement_mapper_and_alphas_bit_exact(gold):
    tok = synth.FakeTokenizer()
    for p in gold["pairs"]:
        m, a = seq_aligner.get_refinement_mapper(p["prompts"], tok)
        assert m.dtype == torch.int64 and m.shape == (1, 77)
        assert m[0].tolist() == p["mapper"]
        assert a[0].tolist() == p["alphas"]


def test_word_indices_equalizer_and_time_gates(gold):
    tok = synth.FakeTokenizer()
    for p in gold["pairs"]:
        src, tgt = p["prompts"]
        assert seq_aligner.get_word_inds(src, p["blend"][0], tok).tolist() == p["inds_src"]
        assert seq_aligner.get_word_inds(tgt, p["blend"][1], tok).tolist() == p["inds_tgt"]
        assert get_equalizer(tgt, (p["eq_word"],), (2,), tok)[0].tolist() == p["equalizer"]
        for n in (50, 20, 3):
            t = ptp_utils.get_time_words_attention_alpha(p["prompts"], n, {"default_": 0.4}, tok)
            assert t.shape == (n + 1, 1, 1, 1, 77)
            assert t.reshape(n + 1, 77)[:, 0].tolist() == p[f"cross_alpha_{n}"]


def test_cross_window_is_20_and_self_window_30_for_50_steps(gold):
    # int(0.4*51) = 20 cross steps vs int(0.6*50) = 30 self steps
    p = gold["pairs"][0]
    assert sum(p["cross_alpha_50"]) == 20


def test_replacement_mapper(gold):
    tok = synth.FakeTokenizer()
    n = 0
    for p in gold["pairs"]:
        if "replace_mapper" in p:
            assert seq_aligner.get_replacement_mapper(p["prompts"], tok)[0].tolist() == p["replace_mapper"]
            n += 1
    assert n >= 2
    with pytest.raises(ValueError):
        seq_aligner.get_replacement_mapper(["a cat", "a big cat"], tok)


def test_schedule_tables(gold):
    s = gold["schedule"]
    sch = scheduler.DDIMSchedulerDev(table_dtype="float64")
    assert float(sch.final_alpha_cumprod) == s["final_alpha_cumprod"]
    assert [float(sch.alphas_cumprod[i]) for i in range(0, 1000, 100)] == s["alphas_cumprod_f64_every_100"]
    for n in (50, 20):  # divisors of 1000: the vendored 0.3.0 scheduler of the fixture and diffusers 0.10 agree
        sch.set_timesteps(n)
        assert sch.timesteps.dtype == torch.int64
        assert sch.timesteps.tolist() == s[f"timesteps_{n}"]
    # non-divisors follow diffusers 0.10 (what DDIMSchedulerDev inherits, models/p2p/scheduler_dev.py:3-10):
    # (arange(0, n) * (1000 // n)).round()[::-1] -> exactly n timesteps (the vendored 0.3.0 arange(0,1000,1000//n)
    # of the fixture yields n + 1 for n = 3 or 30, which overruns noise_loss_list in the forward loops)
    for n, first in ((3, 666), (30, 957), (75, 962), (7, 852)):
        sch.set_timesteps(n)
        ts = sch.timesteps.tolist()
        assert len(ts) == n and ts[0] == first and ts[-1] == 0 and ts == sorted(ts, reverse=True)
        assert ts == [(n - 1 - i) * (1000 // n) for i in range(n)]
    sch.set_timesteps(50)
    assert sch.timesteps.tolist() == list(range(980, -1, -20))
    # the float32 table (diffusers >= 0.10 way) agrees with the float64 one to fp32 precision
    s32 = scheduler.DDIMSchedulerDev(table_dtype="float32")
    assert s32.alphas_cumprod.dtype == torch.float32
    assert abs(float(s32.alphas_cumprod[0]) - 0.99915) < 1e-6
    assert torch.allclose(s32.alphas_cumprod.double(), sch.alphas_cumprod, rtol=2e-5)
