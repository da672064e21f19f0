#!/usr/bin/env python
"""Latency of the device encoders (csrc/encoder_kernels.cuh) next to the same HF modules run by PyTorch on the same GPU.

    python benchmarks/encoder_bench.py > gpurun_out/encoder_bench.json

MiniLM-L6 (384-d), CLIP text tower (512-d), ms-marco cross-encoder; random-init weights of the real architectures
(no checkpoints offline).  Times are CUDA-event means per forward pass after warm-up; `torch_fp32_ms` / `torch_bf16_ms` are
the HF eager modules (what the reference's embeddings.py runs), inputs already on the device."""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
enc_mod = importlib.import_module("multimodal-rag-for-image-text-search_b200.encoders")
native = importlib.import_module("multimodal-rag-for-image-text-search_b200._native")


def timed(fn, reps=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    from transformers import BertConfig, BertForSequenceClassification, BertModel, CLIPTextConfig, CLIPTextModelWithProjection
    torch.manual_seed(0)
    bcfg = dict(vocab_size=30522, hidden_size=384, num_hidden_layers=6, num_attention_heads=12, intermediate_size=1536)
    models = {
        "minilm": (BertModel(BertConfig(**bcfg), add_pooling_layer=False).eval(), "bert"),
        "clip_text": (CLIPTextModelWithProjection(CLIPTextConfig()).eval(), "clip"),
        "cross": (BertForSequenceClassification(BertConfig(**bcfg, num_labels=1)).eval(), "bert"),
    }
    shapes = {"minilm": [(1, 16), (8, 16), (32, 32), (128, 16)], "clip_text": [(1, 16), (8, 16), (128, 16), (32, 77)],
              "cross": [(8, 128), (64, 128), (8, 512)]}
    out = []
    only = os.environ.get("ENC_BENCH_ONLY")              # e.g. "cross": one model family (A/B runs of a switch)
    skip_torch = os.environ.get("ENC_BENCH_SKIP_TORCH") == "1"
    if only == "cross":
        shapes["cross"] = shapes["cross"] + [(16, 256), (128, 128), (32, 512)]
    for name, (model, fam) in models.items():
        if only and name != only:
            continue
        dev = enc_mod.DeviceEncoder.from_hf_bert(model) if fam == "bert" else enc_mod.DeviceEncoder.from_hf_clip(model)
        gpu32 = model.cuda()
        for b, s in shapes[name]:
            ids = np.random.default_rng(b * s).integers(1000, 30000, size=(b, s))
            if fam == "clip":
                ids[:, 0], ids[:, -1] = 49406, 49407
            mask = np.ones_like(ids)
            ids_t, mask_t = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
            n0 = native.lib().mmr_launch_count()
            dev.forward_ids(ids, mask)
            launches = native.lib().mmr_launch_count() - n0
            ours = timed(lambda: dev.forward_ids(ids, mask))
            if skip_torch:
                out.append({"model": name, "batch": b, "seq": s, "device_encoder_ms": ours, "kernel_launches": int(launches)})
                continue
            with torch.no_grad():
                t32 = timed(lambda: gpu32(input_ids=ids_t, attention_mask=mask_t), reps=20)
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    t16 = timed(lambda: gpu32(input_ids=ids_t, attention_mask=mask_t), reps=20)
            out.append({"model": name, "batch": b, "seq": s, "device_encoder_ms": ours, "kernel_launches": int(launches),
                        "torch_fp32_ms": t32, "torch_bf16_autocast_ms": t16, "speedup_vs_torch_fp32": t32 / ours})
        dev.close()
        del gpu32
    print(json.dumps({"what": "ms per forward pass, CUDA events, inputs resident (ours: token ids from host, one 3-array H2D)",
                      "results": out}, indent=1))


if __name__ == "__main__":
    main()
