"""Device-resident feature store + caption batching: the data side of the reference's train / eval loops (dataloader.py:11-53,
consumed at train.py:62-65,116 and eval.py:34-35,48) for a model that eats >30 000 videos/s per GPU.

The reference loads one `.npy` per item and moves it to the device inside `__getitem__` (dataloader.py:37-38).  All of MSVD's
fc7 features are 1970 x 80 x 4096 fp32 = 2.6 GB, so here the whole split lives in ONE device tensor and a batch is an index
gather on the device.  Kept from the reference: the item tuple `(feat, pad_label, ID, mask)` (batched as a DataLoader would
collate it), a caption drawn per item per pass with `np.random.choice(labels, 1)[0]` (dataloader.py:41), truncation / zero padding
to `max_len` and the 1/0 mask (dataloader.py:43-48), and the split filter over `*.npy` stems (dataloader.py:20-24).
"""
from __future__ import annotations

import json
import pathlib as plb
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch


class DeviceFeatureStore:
    def __init__(self, captions_file, feat_path, max_len: int = 80, mode: str = "train", device="cuda", feats_require_grad: bool = False,
                 dtype: torch.dtype = torch.float32):
        """dtype=torch.bfloat16: the features are rounded once, at load time, to what the tensor-core training path would round them to
        at every step anyway (identical loss and gradients; half the memory, no per-step cast).  Only S2VT.forward_loss on the bf16
        training path accepts such batches; decode and the exact path need float32."""
        with open(captions_file, encoding="utf-8") as f:
            data = json.load(f)
        self.word2ix: Dict[str, int] = data["word2ix"]
        self.ix2word: Dict[str, str] = data["ix2word"]
        self.captions: Dict[str, List[List[int]]] = data["captions"]
        self.splits = data["splits"]
        self.max_len = max_len
        self.device = torch.device(device)
        self.feats_require_grad = feats_require_grad          # dataloader.py:38 marks every feature tensor requires_grad=True
        wanted = set(self.splits[mode])
        self.feat_paths = [p for p in plb.Path(feat_path).glob("*.npy") if p.stem in wanted]     # same order as the reference's glob
        self.ids: List[str] = [p.stem for p in self.feat_paths]
        feats = [np.load(str(p)).astype(np.float32, copy=False) for p in self.feat_paths]
        host = torch.from_numpy(np.stack(feats)) if feats else torch.empty(0, max_len, 0)
        if dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("DeviceFeatureStore dtype must be float32 or bfloat16")
        if dtype == torch.bfloat16 and feats_require_grad:
            raise ValueError("bfloat16 features cannot require grad (feats.grad is a float32 quantity of the exact interface)")
        self.feats = host.to(self.device).to(dtype)            # [N, L, F], resident for the whole run

    def __len__(self) -> int:
        return len(self.ids)

    def _labels(self, indices: Sequence[int]) -> Tuple[torch.Tensor, torch.Tensor]:
        pad = np.zeros((len(indices), self.max_len), dtype=np.int64)
        mask = np.zeros((len(indices), self.max_len), dtype=np.float32)
        for r, i in enumerate(indices):
            labels = self.captions[self.ids[i]]
            # dataloader.py:41 calls np.random.choice(labels, 1)[0] on the list of captions; numpy >= 1.24 rejects that ragged list,
            # older numpy turned it into a 1-D object array and drew one index -- the same draw as this:
            label = labels[int(np.random.choice(len(labels), 1)[0])]
            label = label[:self.max_len]
            pad[r, :len(label)] = np.asarray(label, dtype=np.int64)
            mask[r, :len(label)] = 1.0
        return torch.from_numpy(pad).to(self.device, non_blocking=True), torch.from_numpy(mask).to(self.device, non_blocking=True)

    def batch(self, indices: Sequence[int], into=None):
        """(feats [B,L,F], pad_label [B,L] int64, IDs list[str], mask [B,L]) -- what a DataLoader yields over VideoDataset items.
        `into`: a dp.DataParallelTrainer -- features and labels are gathered straight into its static input buffers, so its
        captured train step replays without an extra copy."""
        idx = torch.as_tensor(list(indices), dtype=torch.int64, device=self.device)
        pad, mask = self._labels(indices)
        if into is not None and not self.feats_require_grad and self.feats.is_cuda:
            fb, tb = into.input_buffers((len(indices),) + tuple(self.feats.shape[1:]), tuple(pad.shape), self.feats.dtype)
            torch.index_select(self.feats, 0, idx, out=fb)
            tb.copy_(pad, non_blocking=True)
            return fb, tb, [self.ids[i] for i in indices], mask
        feats = self.feats.index_select(0, idx)
        if self.feats_require_grad:
            feats.requires_grad_(True)
        return feats, pad, [self.ids[i] for i in indices], mask

    def batches(self, batch_size: int, shuffle: bool = False, generator: Optional[torch.Generator] = None, into=None,
                only: Optional[Tuple[int, int, int]] = None) -> Iterator:
        """One pass over the split (train.py:62-65: shuffle=True for training, False for validation / test).
        `only` = (rank, world, usable): yield only batches rank, rank+world, ... of the first `usable` (data-parallel fit())."""
        n = len(self)
        order = torch.randperm(n, generator=generator).tolist() if shuffle else list(range(n))
        for bi, s in enumerate(range(0, n, batch_size)):
            if only is not None and (bi >= only[2] or bi % only[1] != only[0]):
                continue
            yield self.batch(order[s:s + batch_size], into=into)

    def shard(self, rank: int, world: int) -> List[int]:
        """Contiguous item shard of this rank (beam / greedy evaluation shards videos, no collective on the path)."""
        from .dp import shard_range
        lo, hi = shard_range(len(self), rank, world)
        return list(range(lo, hi))


# ------------------------------------------------------------------------------------------------ eval.py post-processing
def ids_to_sentence(pred, ix2word: Dict[str, str], strip_sos: bool = False) -> str:
    """eval.py:54-58 / 90-96: ids -> words, cut at the first '<eos>', drop the first '<sos>' (beam output carries it)."""
    words = [ix2word[str(int(i.item() if hasattr(i, "item") else i))] for i in pred]
    if "<eos>" in words:
        words = words[:words.index("<eos>")]
    if strip_sos and "<sos>" in words:
        words.remove("<sos>")
    return " ".join(words)


def predictions_to_dict(ids: Sequence[str], preds, ix2word: Dict[str, str], beam: bool = False) -> Dict[str, str]:
    """{video id: caption} as eval() / beam_eval() build it (eval.py:46-58, 81-97).  `preds`: int64 [B, L-1] from mode='test', or the
    list[list[Tensor]] from mode='beam_search'."""
    return {ID: ids_to_sentence(p, ix2word, strip_sos=beam) for ID, p in zip(ids, preds)}
