#!/usr/bin/env python
"""A few forward passes of the device encoders for an ncu launch list (per-kernel time shares).

    python benchmarks/encoder_profile.py [cross|minilm|clip] [B] [S]"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
enc_mod = importlib.import_module("multimodal-rag-for-image-text-search_b200.encoders")


def main():
    from transformers import BertConfig, BertForSequenceClassification, BertModel, CLIPTextConfig, CLIPTextModelWithProjection
    kind = sys.argv[1] if len(sys.argv) > 1 else "cross"
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    s = int(sys.argv[3]) if len(sys.argv) > 3 else 128
    torch.manual_seed(0)
    bcfg = dict(vocab_size=30522, hidden_size=384, num_hidden_layers=6, num_attention_heads=12, intermediate_size=1536)
    if kind == "cross":
        enc = enc_mod.DeviceEncoder.from_hf_bert(BertForSequenceClassification(BertConfig(**bcfg, num_labels=1)))
    elif kind == "minilm":
        enc = enc_mod.DeviceEncoder.from_hf_bert(BertModel(BertConfig(**bcfg), add_pooling_layer=False))
    else:
        enc = enc_mod.DeviceEncoder.from_hf_clip(CLIPTextModelWithProjection(CLIPTextConfig()))
    ids = np.random.default_rng(0).integers(1000, 30000, size=(b, s))
    if kind == "clip":
        ids[:, 0], ids[:, -1] = 49406, 49407
    for _ in range(3):
        out = enc.forward_ids(ids, np.ones_like(ids))
    torch.cuda.synchronize()
    print(kind, b, s, float(out.float().abs().sum()))


if __name__ == "__main__":
    main()
