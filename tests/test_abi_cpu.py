"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol declared in
include/s2vt_b200.h, the ctypes signature table covers the header, and the host-side module contract
(constructor, state_dict layout, pickling, error behaviour) matches the reference (S2VTModel.py:11-37)."""
import io
import os
import re

import pytest
import torch

import s2vt_b200
from s2vt_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "s2vt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(s2vt_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    import __graft_entry__ as g
    g.build()
    lib = L.load()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "missing export " + s
        assert s in L.SIGNATURES, "no ctypes signature for " + s
    assert lib.s2vt_abi_version() == 1
    assert lib.s2vt_has_tcgen05() == 1


def test_state_dict_layout_matches_reference():
    V, F, Lq, H, E = 50, 24, 6, 16, 12
    m = s2vt_b200.S2VT(V, F, Lq, dim_hid=H, dim_embed=E)
    sd = m.state_dict()
    expect = {
        "vid_rnn.weight_ih_l0": (4 * H, H), "vid_rnn.weight_hh_l0": (4 * H, H),
        "vid_rnn.bias_ih_l0": (4 * H,), "vid_rnn.bias_hh_l0": (4 * H,),
        "word_rnn.weight_ih_l0": (4 * H, E + H), "word_rnn.weight_hh_l0": (4 * H, H),
        "word_rnn.bias_ih_l0": (4 * H,), "word_rnn.bias_hh_l0": (4 * H,),
        "feat_linear.weight": (H, F), "feat_linear.bias": (H,),
        "out_linear.weight": (V, H), "out_linear.bias": (V,), "embedding.weight": (V, E),
    }
    assert list(sd.keys()) == list(expect.keys())          # same registration order as the reference
    for k, shp in expect.items():
        assert tuple(sd[k].shape) == shp and sd[k].dtype == torch.float32
    assert sum(p.numel() for p in m.parameters()) == sum(torch.Size(s).numel() for s in expect.values())
    for a in ("feat_dim", "length", "dim_hid", "dim_embed", "sos_ix", "eos_ix", "vocab_size", "rnn_type"):
        assert hasattr(m, a)
    m.rnn_type, m.sos_ix, m.eos_ix = "lstm", 3, 4           # eval.py:84-86 overwrites these
    assert s2vt_b200.S2VTModel is s2vt_b200.S2VT


def test_seeded_init_matches_torch_module_init():
    """Same construction order and init calls as the reference => same weights from the same seed."""
    V, F, Lq, H, E = 30, 20, 4, 8, 6
    torch.manual_seed(123)
    m = s2vt_b200.S2VT(V, F, Lq, dim_hid=H, dim_embed=E)
    torch.manual_seed(123)
    vid = torch.nn.LSTM(H, H, batch_first=True)
    word = torch.nn.LSTM(H + E, H, batch_first=True)
    fl = torch.nn.Linear(F, H)
    ol = torch.nn.Linear(H, V)
    emb = torch.nn.Embedding(V, E)
    sd = m.state_dict()
    for name, mod in (("vid_rnn", vid), ("word_rnn", word), ("feat_linear", fl), ("out_linear", ol), ("embedding", emb)):
        for k, v in mod.state_dict().items():
            assert torch.equal(sd[name + "." + k], v), name + "." + k


def test_whole_module_pickle_roundtrip():
    m = s2vt_b200.S2VT(30, 20, 4, dim_hid=8, dim_embed=6)
    buf = io.BytesIO()
    torch.save(m, buf)                                      # train.py:167 saves the whole module
    buf.seek(0)
    m2 = torch.load(buf, weights_only=False)
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k


def test_unsupported_configs_raise():
    with pytest.raises(NotImplementedError):
        s2vt_b200.S2VT(30, 20, 4, rnn_type="gru")
    with pytest.raises(NotImplementedError):
        s2vt_b200.S2VT(30, 20, 4, num_layers=2)
    with pytest.raises(NotImplementedError):
        s2vt_b200.S2VT(30, 20, 4, feat_dropout=0.5)


def test_no_cpu_fallback():
    m = s2vt_b200.S2VT(30, 20, 4, dim_hid=8, dim_embed=6)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(2, 4, 20), targets=torch.zeros(2, 3, dtype=torch.int64), mode="train")


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "s2vt-video-caption_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                assert "oracle" not in open(os.path.join(root, f)).read(), f


def test_att_baseline_state_dict_and_seeded_init():
    """Att_Baseline drop-in: 22 tensors in the reference's registration order (attention_baseline.py:23-33), identical
    weights from the same seed, padding_idx row zeroed, no CPU fallback."""
    V, F, Lq, H, E = 30, 20, 4, 8, 6
    torch.manual_seed(7)
    m = s2vt_b200.Att_Baseline(V, F, Lq, dim_hid=H, dim_embed=E)
    torch.manual_seed(7)
    ref = dict(encoder=torch.nn.LSTM(H, H, batch_first=True, bidirectional=True),
               decoder=torch.nn.LSTM(2 * H + E, H, batch_first=True), feat_linear=torch.nn.Linear(F, H),
               embedding=torch.nn.Embedding(V, E, padding_idx=0), out_linear=torch.nn.Linear(H, V),
               att_enc=torch.nn.Linear(2 * H, H), att_prev_hid=torch.nn.Linear(H, H), att_apply=torch.nn.Linear(H, 1, bias=False))
    sd = m.state_dict()
    keys = [n + "." + k for n, mod in ref.items() for k in mod.state_dict()]
    assert list(sd.keys()) == keys == list(s2vt_b200.ATT_PARAM_ORDER)
    for n, mod in ref.items():
        for k, v in mod.state_dict().items():
            assert torch.equal(sd[n + "." + k], v), n + "." + k
    assert not sd["embedding.weight"][0].any()
    for a in ("dim_feat", "length", "dim_hid", "dim_embed", "sos_ix", "eos_ix", "vocab_size"):
        assert hasattr(m, a)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(2, 4, 20), targets=torch.zeros(2, 3, dtype=torch.int64), mode="train")
    with pytest.raises(NotImplementedError):
        s2vt_b200.Att_Baseline(V, F, Lq, feat_dropout=0.1)
