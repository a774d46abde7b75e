"""End-to-end slice of the reference's train() / eval() recipe on the sm_100a path: feature store -> fit() (FusedAdam,
ReduceLROnPlateau, EarlyStopping, whole-module checkpoints) -> torch.load -> greedy decode -> captions (train.py:56-168,
eval.py:30-60).  Needs a B200 (-m gpu)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import s2vt_b200  # noqa: E402


def test_fit_checkpoint_reload_and_caption(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    rng = np.random.default_rng(0)
    V, F, Lq = 40, 32, 8
    fd = tmp_path / "feats"
    fd.mkdir()
    ids = ["v%d" % i for i in range(24)]
    caps = {}
    for i in ids:
        np.save(fd / (i + ".npy"), rng.standard_normal((Lq, F)).astype(np.float32))
        caps[i] = [[3] + [int(x) for x in rng.integers(5, V, size=4)] + [4]]
    ix2word = {str(i): "w%d" % i for i in range(V)}
    ix2word.update({"0": "<pad>", "3": "<sos>", "4": "<eos>"})
    data = {"word2ix": {w: int(i) for i, w in ix2word.items()}, "ix2word": ix2word, "captions": caps,
            "splits": {"train": ids[:16], "valid": ids[16:20], "test": ids[20:]}}
    cf = tmp_path / "captions.json"
    cf.write_text(json.dumps(data))
    train = s2vt_b200.DeviceFeatureStore(str(cf), str(fd), max_len=Lq, mode="train")
    valid = s2vt_b200.DeviceFeatureStore(str(cf), str(fd), max_len=Lq, mode="valid")
    test = s2vt_b200.DeviceFeatureStore(str(cf), str(fd), max_len=Lq, mode="test")
    torch.manual_seed(0)
    np.random.seed(0)
    model = s2vt_b200.S2VT(V, F, Lq, dim_hid=32, dim_embed=16).cuda()
    hist = s2vt_b200.fit(model, train, valid, epochs=6, batch_size=8, lr=5e-3, save_path=str(tmp_path / "ckpt"), tag="t_")
    assert len(hist) == 6 and hist[-1]["train_loss"] < hist[0]["train_loss"]
    assert os.path.exists(tmp_path / "ckpt" / "t_final.pth") and os.path.exists(tmp_path / "ckpt" / "t_stop.pth")
    loaded = torch.load(tmp_path / "ckpt" / "t_final.pth", weights_only=False).cuda()      # eval.py:41 loads the whole module
    feats, _, vid_ids, _ = test.batch(list(range(len(test))))
    a = model(feats, mode="test")
    b = loaded(feats, mode="test")
    assert torch.equal(a, b)
    pred = s2vt_b200.predictions_to_dict(vid_ids, b, test.ix2word)
    assert set(pred) == set(ids[20:]) and all(isinstance(s, str) for s in pred.values())
