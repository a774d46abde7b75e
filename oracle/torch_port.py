"""CPU PyTorch port of the reference's S2VT train step / greedy decode -- TEST & BASELINE INFRASTRUCTURE.

The reference's hot path is a Python module that delegates all arithmetic to torch library calls
(nn.LSTM -> oneDNN mkldnn_rnn_layer on CPU, nn.Linear -> MKL, CrossEntropyLoss, optim.Adam;
S2VTModel.py:19-28, utils.py:11, train.py:89).  /root/reference does not exist on the GPU box, so
bench.py's `cpu_baseline` and `--impl reference` legs time THIS port instead: the same library calls in the
same order on the same shapes, written from the behavioural spec in SURVEY.md section 3 (kind = "port").
It is pinned by tests/test_oracle_golden.py::test_torch_port_matches_golden against outputs of the real
reference.  Only tests/ and bench.py may import this file.
"""
from __future__ import annotations

import torch
from torch import nn
import torch.nn.functional as F


class S2VTCpuPort(nn.Module):
    def __init__(self, V: int, feat_dim: int, length: int, H: int, E: int, sos_ix: int = 3):
        super().__init__()
        self.vid_rnn = nn.LSTM(H, H, batch_first=True)
        self.word_rnn = nn.LSTM(H + E, H, batch_first=True)
        self.feat_linear = nn.Linear(feat_dim, H)
        self.out_linear = nn.Linear(H, V)
        self.embedding = nn.Embedding(V, E)
        self.L, self.H, self.E, self.V, self.sos_ix = length, H, E, V, sos_ix

    def _encode(self, feats):
        B = feats.shape[0]
        x = self.feat_linear(feats)                                                  # S2VTModel.py:54
        x = torch.cat([x, x.new_zeros(B, self.L - 1, self.H)], dim=1)                # S2VTModel.py:64-65
        out1, _ = self.vid_rnn(x)                                                    # S2VTModel.py:67
        return out1

    def train_logits(self, feats, targets_in):
        B = feats.shape[0]
        out1 = self._encode(feats)
        emb = self.embedding(targets_in)                                             # S2VTModel.py:71
        emb = torch.cat([emb.new_zeros(B, self.L, self.E), emb], dim=1)              # S2VTModel.py:72-73
        out2, _ = self.word_rnn(torch.cat([emb, out1], dim=2))                       # S2VTModel.py:75-77
        return self.out_linear(out2[:, self.L:, :])                                  # S2VTModel.py:78-80

    @staticmethod
    def criterion(logits, targets, mask):
        """utils.py:13-26: scalar mean CE, then the (cancelling) mask algebra."""
        n = logits.shape[0] * logits.shape[1]
        loss = F.cross_entropy(logits.reshape(n, -1), targets[:, 1:].reshape(-1))
        m = mask[:, 1:].reshape(-1)
        return torch.sum(loss * m) / torch.sum(m)

    @torch.no_grad()
    def greedy(self, feats):
        B = feats.shape[0]
        out1 = self._encode(feats)
        enc_in = torch.cat([out1.new_zeros(B, self.L, self.E), out1[:, :self.L]], dim=2)
        _, st = self.word_rnn(enc_in)                                                # S2VTModel.py:84-86
        tok = torch.full((B,), self.sos_ix, dtype=torch.long)
        pred = []
        for k in range(self.L - 1):                                                  # S2VTModel.py:89-107
            x = torch.cat([self.embedding(tok).unsqueeze(1), out1[:, self.L + k].unsqueeze(1)], dim=2)
            o, st = self.word_rnn(x, st)
            tok = self.out_linear(o.squeeze(1)).argmax(dim=1)
            pred.append(tok)
        return torch.stack(pred, dim=1)

    def train_step(self, opt, feats, targets, mask):
        """train.py:116-127: zero_grad, forward, criterion, backward, Adam step, loss.item()."""
        opt.zero_grad()
        feats = feats.detach().requires_grad_(True)                                  # dataloader.py:38
        loss = self.criterion(self.train_logits(feats, targets[:, :-1]), targets, mask)
        loss.backward()
        opt.step()
        return loss.item()
