"""CPU tests of the host-side rows next to the hot path: the device-resident feature store (N1), eval.py's
id -> caption post-processing (N2), EarlyStopping (N3) and GloVe loading (N4).  Where the reference's own code runs in this
container it is the expectation; elsewhere the reference's documented behaviour is."""
import json
import os
import sys

import numpy as np
import pytest
import torch

import s2vt_b200

REF = "/root/reference"


@pytest.fixture()
def tiny_dataset(tmp_path):
    rng = np.random.default_rng(3)
    feats_dir = tmp_path / "feats"
    feats_dir.mkdir()
    ids = ["vid%d" % i for i in range(7)]
    for i in ids:
        np.save(feats_dir / (i + ".npy"), rng.standard_normal((6, 5)).astype(np.float32))
    caps = {i: [[3] + list(rng.integers(5, 20, size=int(rng.integers(1, 9)))) + [4] for _ in range(int(rng.integers(1, 4)))] for i in ids}
    caps = {k: [[int(x) for x in c] for c in v] for k, v in caps.items()}
    data = {"word2ix": {"<pad>": 0, "<sos>": 3, "<eos>": 4}, "ix2word": {str(i): "w%d" % i for i in range(20)},
            "captions": caps, "splits": {"train": ids[:4], "valid": ids[4:6], "test": ids[6:]}}
    data["ix2word"].update({"0": "<pad>", "3": "<sos>", "4": "<eos>"})
    cf = tmp_path / "captions.json"
    cf.write_text(json.dumps(data))
    return str(cf), str(feats_dir), data


def test_feature_store_tuple_padding_mask_and_split(tiny_dataset):
    cf, fd, data = tiny_dataset
    st = s2vt_b200.DeviceFeatureStore(cf, fd, max_len=6, mode="train", device="cpu")
    assert len(st) == 4 and sorted(st.ids) == sorted(data["splits"]["train"])
    np.random.seed(5)
    feats, pad, ids, mask = st.batch([0, 2, 3])
    assert feats.shape == (3, 6, 5) and feats.dtype == torch.float32
    assert pad.shape == (3, 6) and pad.dtype == torch.int64 and mask.shape == (3, 6) and mask.dtype == torch.float32
    np.random.seed(5)                                       # same draws as dataloader.py:41 makes per item
    for r, i in enumerate([0, 2, 3]):
        labels = data["captions"][st.ids[i]]
        lab = labels[int(np.random.choice(len(labels), 1)[0])][:6]
        assert pad[r, :len(lab)].tolist() == lab and not pad[r, len(lab):].any()
        assert mask[r].tolist() == [1.0] * len(lab) + [0.0] * (6 - len(lab))
        assert np.array_equal(feats[r].numpy(), np.load(os.path.join(fd, st.ids[i] + ".npy")))
    assert ids == [st.ids[i] for i in [0, 2, 3]]
    seen = [i for _, _, b, _ in st.batches(3, shuffle=True, generator=torch.Generator().manual_seed(1)) for i in b]
    assert sorted(seen) == sorted(st.ids)
    assert st.shard(0, 2) + st.shard(1, 2) == list(range(4))


def test_ids_to_sentence_matches_eval_py():
    ix2word = {"0": "<pad>", "3": "<sos>", "4": "<eos>", "7": "a", "8": "cat", "9": "runs"}
    greedy = torch.tensor([[7, 8, 9, 4, 0, 0], [8, 8, 8, 8, 8, 8]])
    # eval.py:54-58
    expect = []
    for pred in greedy:
        words = [ix2word[str(i.item())] for i in pred]
        if "<eos>" in words:
            words = words[:words.index("<eos>")]
        expect.append(" ".join(words))
    assert [s2vt_b200.ids_to_sentence(p, ix2word) for p in greedy] == expect == ["a cat runs", "cat cat cat cat cat cat"]
    beam = [[torch.tensor([[3]]), torch.tensor(7), torch.tensor(8), torch.tensor(4)]]      # <sos> first (S2VTModel.py:231-238)
    assert s2vt_b200.predictions_to_dict(["v"], beam, ix2word, beam=True) == {"v": "a cat"}


def test_early_stopping_matches_reference_semantics(tmp_path):
    """Same decisions as utils.py's EarlyStopping (patched only for numpy 2's removal of np.Inf, utils.py:52)."""
    np.Inf = np.inf                                            # the reference needs this alias to import-run on numpy 2
    sys.path.insert(0, REF)
    try:
        from utils import EarlyStopping as RefES
    except Exception:
        pytest.skip("reference utils.py not importable here")
    finally:
        sys.path.remove(REF)
    losses = [1.0, 0.9, 0.95, 0.91, 0.85, 0.86, 0.87, 0.88]
    ours = s2vt_b200.EarlyStopping(patience=3, path=str(tmp_path / "a.pt"), trace_func=lambda *_: None)
    ref = RefES(patience=3, path=str(tmp_path / "b.pt"), trace_func=lambda *_: None)
    model = torch.nn.Linear(2, 2)
    for l in losses:
        ours(l, model)
        ref(l, model)
        assert (ours.counter, ours.early_stop, ours.best_score, ours.val_loss_min) == (ref.counter, ref.early_stop, ref.best_score, ref.val_loss_min)
    assert ours.early_stop and os.path.exists(tmp_path / "a.pt")


def test_load_glove_weights(tmp_path):
    m = s2vt_b200.S2VT(12, 8, 4, dim_hid=8, dim_embed=4)
    glove = tmp_path / "glove.txt"
    glove.write_text("cat 1 2 3 4\ndog 5 6 7 8\nzebra 9 9 9 9\n")
    ix2word = {"5": "cat", "6": "dog", "7": "unknownword"}
    torch.manual_seed(0)
    n = m.load_glove_weights(str(glove), 4, ix2word)
    assert n == 2
    assert m.embedding.weight[5].tolist() == [1, 2, 3, 4] and m.embedding.weight[6].tolist() == [5, 6, 7, 8]
    assert m.embedding.weight.requires_grad and m.embedding.weight[7].abs().sum() > 0          # xavier-normal row, still trainable
    assert list(m.state_dict().keys())[-1] == "embedding.weight"


def test_wavefront_chunk_boundaries_are_valid_for_every_length():
    """engine_bf16._time_bounds: strictly increasing from 0 to T = 2L - 1, step L is always a boundary (the embedding half of word_rnn's
    input starts there), at most S2VT_MAX_SYNC chunks -- forwards and for BPTT, for every caption length the wave front accepts."""
    from s2vt_b200 import engine_bf16 as EB
    for Lq in range(16, 400):
        T = 2 * Lq - 1
        for backward in (False, True):
            b = EB._time_bounds(Lq, T, backward)
            assert b[0] == 0 and b[-1] == T and Lq in b
            assert all(x < y for x, y in zip(b, b[1:]))
            assert len(b) - 1 <= EB.MAX_SYNC


def test_max_sync_matches_header():
    import re
    from s2vt_b200 import engine_bf16 as EB
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "s2vt_b200.h")).read()
    assert int(re.search(r"#define\s+S2VT_MAX_SYNC\s+(\d+)", hdr).group(1)) == EB.MAX_SYNC


def test_feature_store_bf16_holds_the_rounded_features(tiny_dataset):
    cf, fd, data = tiny_dataset
    st32 = s2vt_b200.DeviceFeatureStore(cf, fd, max_len=6, mode="train", device="cpu")
    st16 = s2vt_b200.DeviceFeatureStore(cf, fd, max_len=6, mode="train", device="cpu", dtype=torch.bfloat16)
    assert st16.feats.dtype == torch.bfloat16 and st16.ids == st32.ids
    assert torch.equal(st16.feats, st32.feats.to(torch.bfloat16))            # round-to-nearest-even, once, at load time
    np.random.seed(1)
    f16 = st16.batch([1, 3])[0]
    assert f16.dtype == torch.bfloat16 and torch.equal(f16, st32.feats[[1, 3]].to(torch.bfloat16))
    with pytest.raises(ValueError):
        s2vt_b200.DeviceFeatureStore(cf, fd, max_len=6, mode="train", device="cpu", dtype=torch.float16)
    with pytest.raises(ValueError):
        s2vt_b200.DeviceFeatureStore(cf, fd, max_len=6, mode="train", device="cpu", dtype=torch.bfloat16, feats_require_grad=True)


def test_profiler_detection_and_numa_binding_degrade_gracefully(monkeypatch):
    from s2vt_b200 import dp, ops
    for k in ("NV_COMPUTE_PROFILER_PERFWORKS_DIR", "CUDA_INJECTION64_PATH"):
        monkeypatch.delenv(k, raising=False)
    assert not ops.serialising_profiler_attached()
    monkeypatch.setenv("CUDA_INJECTION64_PATH", "/opt/nvidia/nsight-compute/2025.2.1/target/linux-desktop-glibc_2_11_3-x64/libcuda-injection.so")
    assert ops.serialising_profiler_attached()
    monkeypatch.setenv("CUDA_INJECTION64_PATH", "/opt/nsys/libToolsInjection64.so")        # a tracer, not a serialising profiler
    assert not ops.serialising_profiler_attached()
    before = os.sched_getaffinity(0)
    assert dp.bind_to_local_numa(0) is None or isinstance(dp.bind_to_local_numa(0), int)  # no GPU / no topology: None, never raises
    assert os.sched_getaffinity(0) <= before
    os.sched_setaffinity(0, before)


def test_rank_sharded_batches_partition_the_epoch(tiny_dataset):
    """fit() under torch.distributed: every rank draws the same permutation and takes every world-th batch of the first `usable`
    ones -- the ranks' batches are disjoint, together they are the first `usable` batches, and every rank gets the same count."""
    cf, fd, data = tiny_dataset
    st = s2vt_b200.DeviceFeatureStore(cf, fd, max_len=6, mode="train", device="cpu")
    bs, world = 1, 3
    n_batches = (len(st) + bs - 1) // bs
    usable = n_batches - n_batches % world
    full = [b for _, _, b, _ in st.batches(bs, shuffle=True, generator=torch.Generator().manual_seed(7))]
    per_rank = []
    for r in range(world):
        per_rank.append([b for _, _, b, _ in st.batches(bs, shuffle=True, generator=torch.Generator().manual_seed(7), only=(r, world, usable))])
    assert len({len(x) for x in per_rank}) == 1 and len(per_rank[0]) == usable // world
    merged = [per_rank[i % world][i // world] for i in range(usable)]
    assert merged == full[:usable]


def test_fused_adam_rejects_param_groups_and_keeps_flat_state_keys():
    a, b = torch.nn.Parameter(torch.zeros(8)), torch.nn.Parameter(torch.zeros(8))
    with pytest.raises(ValueError):
        s2vt_b200.FusedAdam([{"params": [a]}, {"params": [b], "lr": 1e-2}])
    opt = s2vt_b200.FusedAdam([a, b], lr=1e-3)
    sd = opt.state_dict()                                 # before the first step there is no flat state yet: plain torch layout
    assert "fused" not in sd and "param_groups" in sd
