"""Device encoders against plain fp32 PyTorch on the SAME weights.

The HF modules are built from their configs with seeded random weights (no checkpoint files offline); the linear weights
are rounded to bf16-representable values first, so both sides hold identical weights and only the activation precision
differs (bf16 GEMM inputs on the device, fp32 in torch).  Tolerance: |delta| <= 1e-3 per component of
the unit-norm embedding; the cross-encoder's raw logit is not normalised, so its bound is relative: 2e-3 + 1e-2 |logit|,
and the ORDER of the candidates (what _rerank_text uses) must agree wherever torch's logits differ by more than that."""
import importlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
PKG = "multimodal-rag-for-image-text-search_b200"
TOL = 1e-3


def _round_linear_weights(model):
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() == 2 and "embedding" not in name.lower():
                p.copy_(p.to(torch.bfloat16).to(torch.float32))
    return model.eval()


def _spread(model, scale=2.0):
    """Random init (std 0.02) gives near-uniform attention; scale the projections up so softmax / GELU / LN see spread."""
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() == 2 and "embedding" not in name.lower():
                p.mul_(scale)
            elif p.dim() == 1 and "bias" in name:
                p.normal_(0.0, 0.05)
            elif p.dim() == 1 and "weight" in name:
                p.normal_(1.0, 0.1)
    return model


@pytest.fixture(scope="module")
def enc_mod():
    return importlib.import_module(PKG + ".encoders")


@pytest.fixture(scope="module")
def bert():
    from transformers import BertConfig, BertModel
    torch.manual_seed(0)
    cfg = BertConfig(vocab_size=30522, hidden_size=384, num_hidden_layers=6, num_attention_heads=12, intermediate_size=1536,
                     max_position_embeddings=512)
    return _round_linear_weights(_spread(BertModel(cfg, add_pooling_layer=False)))


def _bert_reference(model, ids, mask, types=None):
    with torch.no_grad():
        h = model(input_ids=ids, attention_mask=mask, token_type_ids=types).last_hidden_state
        m = mask[..., None].float()
        e = (h * m).sum(1) / m.sum(1).clamp(min=1e-9)          # sentence-transformers mean pooling
        return torch.nn.functional.normalize(e, dim=1)         # Normalize module + the reference's _normalize


@pytest.mark.parametrize("b,s", [(1, 12), (5, 40), (33, 7), (2, 130), (128, 16)])
def test_minilm_text_encoder_matches_fp32_torch(enc_mod, bert, b, s):
    g = torch.Generator().manual_seed(b * 1000 + s)
    ids = torch.randint(1000, 30000, (b, s), generator=g)
    lens = torch.randint(max(1, s // 3), s + 1, (b,), generator=g)
    lens[0] = s
    mask = (torch.arange(s)[None, :] < lens[:, None]).long()
    enc = enc_mod.DeviceEncoder.from_hf_bert(bert)
    assert enc.kind == "minilm" and enc.out_dim == 384
    got = enc.forward_ids(ids.numpy(), mask.numpy()).cpu()
    want = _bert_reference(bert, ids, mask)
    assert got.shape == want.shape
    assert (got - want).abs().max().item() <= TOL, (got - want).abs().max().item()
    assert torch.allclose(got.norm(dim=1), torch.ones(b), atol=1e-5)
    # mask = None means "no padding"
    got2 = enc.forward_ids(ids.numpy()).cpu()
    assert (got2 - _bert_reference(bert, ids, torch.ones_like(ids))).abs().max().item() <= TOL
    enc.close()


def test_layernorm_folded_into_the_gemm_matches_the_separate_kernel(enc_mod, bert):
    """gemm_wt_kernel<EPI, 64, true> computes the LayerNorm in front of a GEMM itself (shipped for passes of <= 16 tokens,
    MMR_ENC_FUSE_LN=2 forces it for every pass on the 64-token tile): bit-identical embeddings to the layernorm_kernel path
    (a request's embedding must not depend on how many tokens shared its pass), inside the fp32-torch tolerance."""
    native = importlib.import_module(PKG + "._native")
    enc = enc_mod.DeviceEncoder.from_hf_bert(bert)
    g = torch.Generator().manual_seed(77)
    try:
        for b, s in ((1, 12), (1, 16), (4, 40), (9, 100)):     # 12 / 16 / 160 / 900 tokens: 1, 1, 3 and 15 token tiles
            ids = torch.randint(1000, 30000, (b, s), generator=g)
            lens = torch.randint(max(1, s // 2), s + 1, (b,), generator=g)
            lens[0] = s
            mask = (torch.arange(s)[None, :] < lens[:, None]).long()
            want = _bert_reference(bert, ids, mask)
            outs = {}
            for mode in ("0", "2"):
                native.set_option("MMR_ENC_FUSE_LN", mode)
                n0 = native.lib().mmr_launch_count()
                outs[mode] = enc.forward_ids(ids.numpy(), mask.numpy()).cpu()
                outs[mode + "n"] = native.lib().mmr_launch_count() - n0
                assert (outs[mode] - want).abs().max().item() <= TOL, (mode, b, s)
            assert outs["2n"] == outs["0n"] - 11, (outs["0n"], outs["2n"])       # 2 per layer, minus the last layer's closing LN
            assert torch.equal(outs["0"], outs["2"]), (b, s)       # same arithmetic, bit for bit: batch shape cannot change an answer
    finally:
        native.set_option("MMR_ENC_FUSE_LN", None)
        enc.close()


def test_clip_text_tower_matches_fp32_torch(enc_mod):
    from transformers import CLIPTextConfig, CLIPTextModelWithProjection
    torch.manual_seed(1)
    model = _round_linear_weights(_spread(CLIPTextModelWithProjection(CLIPTextConfig()), 2.0))
    enc = enc_mod.DeviceEncoder.from_hf_clip(model)
    assert enc.kind == "clip_text" and enc.out_dim == 512
    for b, s in ((1, 9), (6, 25), (40, 77)):
        g = torch.Generator().manual_seed(s)
        ids = torch.randint(1000, 40000, (b, s), generator=g)
        ids[:, 0] = 49406
        lens = torch.randint(3, s + 1, (b,), generator=g)
        lens[0] = s
        mask = (torch.arange(s)[None, :] < lens[:, None]).long()
        for i in range(b):                                   # EOS closes every sequence, padding repeats it (CLIP tokenizer)
            ids[i, lens[i] - 1:] = 49407
        with torch.no_grad():
            want = torch.nn.functional.normalize(model(input_ids=ids, attention_mask=mask).text_embeds, dim=1)
        got = enc.forward_ids(ids.numpy(), mask.numpy()).cpu()
        assert (got - want).abs().max().item() <= TOL, (b, s, (got - want).abs().max().item())
    enc.close()


def test_cross_encoder_logits_match_fp32_torch(enc_mod):
    from transformers import BertConfig, BertForSequenceClassification
    torch.manual_seed(2)
    cfg = BertConfig(vocab_size=30522, hidden_size=384, num_hidden_layers=6, num_attention_heads=12, intermediate_size=1536,
                     max_position_embeddings=512, num_labels=1)
    model = _round_linear_weights(_spread(BertForSequenceClassification(cfg)))
    enc = enc_mod.DeviceEncoder.from_hf_bert(model)
    assert enc.kind == "cross" and enc.out_dim == 1
    for b, s in ((8, 64), (3, 300), (16, 128), (72, 64)):      # token tiles of 64, 128 and 256
        g = torch.Generator().manual_seed(s)
        ids = torch.randint(1000, 30000, (b, s), generator=g)
        lens = torch.randint(s // 2, s + 1, (b,), generator=g)
        mask = (torch.arange(s)[None, :] < lens[:, None]).long()
        types = (torch.arange(s)[None, :] >= (lens[:, None] // 3)).long() * mask      # query segment 0, passage segment 1
        with torch.no_grad():
            want = model(input_ids=ids, attention_mask=mask, token_type_ids=types).logits[:, 0]
        got = enc.forward_ids(ids.numpy(), mask.numpy(), types.numpy()).cpu()
        assert got.shape == (b,)
        bound = 2e-3 + 1e-2 * want.abs()
        assert bool(((got - want).abs() <= bound).all()), (b, s, (got - want).abs().max().item(), want[:4], got[:4])
        for i in range(b):
            for j in range(b):
                if want[i] - want[j] > 2 * bound.max():
                    assert got[i] > got[j], "rerank order differs from fp32 torch beyond the tolerance"
    enc.close()


def test_cross_encoder_logit_does_not_depend_on_the_batch_it_rides_in(enc_mod):
    """attention_mma_kernel (tensor-core attention of the cross-encoder): a pair's logit is bit-identical whether it is scored
    alone, padded to a longer neighbour, or inside a large micro-batch (masked / zero-filled keys add exact zeros; the kernel
    is chosen by model kind, never by batch shape)."""
    from transformers import BertConfig, BertForSequenceClassification
    torch.manual_seed(5)
    cfg = BertConfig(vocab_size=30522, hidden_size=384, num_hidden_layers=6, num_attention_heads=12, intermediate_size=1536,
                     max_position_embeddings=512, num_labels=1)
    enc = enc_mod.DeviceEncoder.from_hf_bert(_round_linear_weights(_spread(BertForSequenceClassification(cfg))))
    g = torch.Generator().manual_seed(11)
    lens = [40, 300, 17, 129, 64, 512, 96, 200]
    seqs = [torch.randint(1000, 30000, (n,), generator=g) for n in lens]

    def run(idx):
        s = max(lens[i] for i in idx)
        ids = torch.zeros((len(idx), s), dtype=torch.long)
        mask = torch.zeros((len(idx), s), dtype=torch.long)
        types = torch.zeros((len(idx), s), dtype=torch.long)
        for r, i in enumerate(idx):
            ids[r, : lens[i]], mask[r, : lens[i]] = seqs[i], 1
            types[r, lens[i] // 3: lens[i]] = 1
        return enc.forward_ids(ids.numpy(), mask.numpy(), types.numpy()).cpu()

    together = run(list(range(len(lens))))
    for i in range(len(lens)):
        alone = run([i])
        assert torch.equal(alone[0], together[i]), (i, lens[i], alone[0].item(), together[i].item())
    pair = run([0, 1])
    assert torch.equal(pair[0], together[0]) and torch.equal(pair[1], together[1])
    many = run([2, 0] * 40)                    # 80 x 40 = 3200 tokens: another token tile of the GEMMs
    assert torch.equal(many[1], together[0]) and torch.equal(many[0], together[2])
    enc.close()


def test_query_is_born_on_the_device_and_feeds_the_scan(enc_mod, bert):
    """embed -> search without a host hop: the encoder's output tensor IS mmr_search's query buffer."""
    from tests import util
    mmr = importlib.import_module(PKG)
    rows = util.unit_rows(50_000, 384, seed=5)
    ix = mmr.ResidentIndex.from_f32(rows, dtype="bf16")
    enc = enc_mod.DeviceEncoder.from_hf_bert(bert)

    class Tok:  # a stand-in tokenizer (the real one needs vocab files): hashes words to ids
        def __call__(self, texts, padding=True, truncation=True, return_tensors="np", max_length=256):
            toks = [[101] + [1000 + (hash(w) % 20000) for w in t.split()][: max_length - 2] + [102] for t in texts]
            s = max(len(t) for t in toks)
            ids = np.zeros((len(toks), s), np.int64)
            mask = np.zeros((len(toks), s), np.int64)
            for i, t in enumerate(toks):
                ids[i, : len(t)], mask[i, : len(t)] = t, 1
            return {"input_ids": ids, "attention_mask": mask, "token_type_ids": np.zeros_like(ids)}

    text_enc = enc_mod.TextQueryEncoder(Tok(), enc)
    q_dev = text_enc.encode_device(["what is in the picture", "a much longer question about retrieval augmented generation"])
    assert q_dev.is_cuda and q_dev.shape == (2, 384)
    s, r = ix.search(q_dev, 10)
    host = text_enc(["what is in the picture", "a much longer question about retrieval augmented generation"])
    s2, r2 = ix.search_host(host, 10)
    assert (r.cpu().numpy() == r2).all()
    for j in range(2):
        util.check_topk(s[j].cpu().numpy(), r[j].cpu().numpy(), util.oracle_scores(rows, host[j]), 10, util.TOL_BF16, what="enc->scan")
    assert text_enc([]).shape == (0, 384)
    enc.close()
    ix.close()


class _Tok:
    """Stand-in tokenizer (the real ones need vocabulary files): words hash to ids; pairs get token types 0 / 1."""

    def __init__(self, bos=101, eos=102, lo=1000, span=20000, clip=False):
        self.bos, self.eos, self.lo, self.span, self.clip = bos, eos, lo, span, clip

    def _ids(self, text):
        import zlib
        return [self.lo + (zlib.crc32(w.encode()) % self.span) for w in text.split()]

    def __call__(self, texts, pairs=None, padding=True, truncation=True, return_tensors="np", max_length=256):
        seqs, types = [], []
        for n, t in enumerate(texts):
            a = [self.bos] + self._ids(t)[: max_length - 2] + [self.eos]
            ty = [0] * len(a)
            if pairs is not None:
                b = self._ids(pairs[n])[: max(0, max_length - len(a) - 1)] + [self.eos]
                a, ty = a + b, ty + [1] * len(b)
            seqs.append(a)
            types.append(ty)
        s = max(len(a) for a in seqs)
        pad = self.eos if self.clip else 0
        ids = np.full((len(seqs), s), pad, np.int64)
        mask = np.zeros((len(seqs), s), np.int64)
        tt = np.zeros((len(seqs), s), np.int64)
        for i, (a, ty) in enumerate(zip(seqs, types)):
            ids[i, : len(a)], mask[i, : len(a)], tt[i, : len(a)] = a, 1, ty
        return {"input_ids": ids, "attention_mask": mask, "token_type_ids": tt}


def test_whole_request_path_on_the_device_equals_per_request_host_path(enc_mod, bert):
    """SURVEY 8f rank 2 + 3 end to end: B requests -> MiniLM + CLIP text tower on the device -> two scans -> ONE cross-encoder
    pass over all (query, passage) pairs -> mmr_fuse_f64 (rerank re-ordering + z-score fusion + CONFIDENCE_TAU gate).
    Must equal, request by request, the reference-shaped host path `retrieve()` + `_confidence_low()` driven by the same
    models: same winners, bit-identical combined scores, same gate."""
    from types import SimpleNamespace
    from transformers import BertConfig, BertForSequenceClassification, CLIPTextConfig, CLIPTextModelWithProjection
    from tests import util
    mmr = importlib.import_module(PKG)
    retrieve = importlib.import_module(PKG + ".retrieve")
    cache = importlib.import_module(PKG + ".cache")
    settings_mod = importlib.import_module(PKG + ".settings")
    torch.manual_seed(7)
    clip = _round_linear_weights(_spread(CLIPTextModelWithProjection(CLIPTextConfig(num_hidden_layers=4)), 2.0))
    cross = _round_linear_weights(_spread(BertForSequenceClassification(BertConfig(
        vocab_size=30522, hidden_size=384, num_hidden_layers=6, num_attention_heads=12, intermediate_size=1536, num_labels=1))))
    text_enc = enc_mod.TextQueryEncoder(_Tok(), enc_mod.DeviceEncoder.from_hf_bert(bert))
    image_enc = enc_mod.ImageQueryEncoder(_Tok(bos=49406, eos=49407, span=40000, clip=True), enc_mod.DeviceEncoder.from_hf_clip(clip))
    tok = _Tok()
    cross_enc = enc_mod.DeviceCrossEncoder(lambda a, b, **kw: tok(a, pairs=b, max_length=kw.get("max_length", 512)),
                                           enc_mod.DeviceEncoder.from_hf_bert(cross))
    n_t, n_i = 6000, 2500
    users = ["u0", "u1", "u2"]
    store = mmr.B200Store()
    store.load_arrow("text_collection", mmr.make_arrow_table([f"t{i}" for i in range(n_t)], [users[i % 3] for i in range(n_t)],
                                                             ["d"] * n_t, ["text"] * n_t, util.unit_rows(n_t, 384, 71), ["{}"] * n_t))
    store.load_arrow("image_collection", mmr.make_arrow_table([f"i{i}" for i in range(n_i)], [users[i % 3] for i in range(n_i)],
                                                              ["d"] * n_i, ["image"] * n_i, util.unit_rows(n_i, 512, 72), ["{}"] * n_i))
    chunks = {f"t{i}": SimpleNamespace(id=f"t{i}", document_id="d", modality="text", text=f"passage number {i} about topic {i % 17}",
                                       meta={}, page_no=i, start_ts=None, end_ts=None, file_path=None) for i in range(n_t)}
    chunks.update({f"i{i}": SimpleNamespace(id=f"i{i}", document_id="d", modality="image", text=None, meta={}, page_no=None,
                                            start_ts=None, end_ts=None, file_path=f"/f/{i}.jpg") for i in range(n_i)})
    meta = SimpleNamespace(get_chunk=chunks.get, get_chunks=lambda ids: {c: chunks[c] for c in ids if c in chunks})
    cache.clear_all_caches()
    retrieve.configure(store=store, metadata=meta, device_text_encoder=text_enc, device_image_encoder=image_enc,
                       device_cross_encoder=cross_enc, retrieval_settings=settings_mod.RetrievalSettings(use_rerank=True))
    queries = [f"question {j} about topic {j % 5} and retrieval" for j in range(7)]
    who = [users[j % 3] for j in range(7)]
    got = retrieve.retrieve_batch_device(who, queries)
    for j in range(7):
        cache.clear_all_caches()
        host = retrieve.retrieve(who[j], queries[j])
        items, low = got[j]
        assert [it["chunk_id"] for it in items] == [h["chunk_id"] for h in host], j
        assert [it["combined_score"] for it in items] == [h["combined_score"] for h in host], j
        assert [it.get("rerank_score") for it in items] == [h.get("rerank_score") for h in host], j
        assert low is retrieve._confidence_low(host)
        assert len(items) == 4 and all("metadata" in it for it in items)
    # rerank off: same comparison through mmr_fuse
    retrieve.configure(retrieval_settings=settings_mod.RetrievalSettings(use_rerank=False))
    cache.clear_all_caches()
    got = retrieve.retrieve_batch_device(who, queries)
    for j in range(7):
        cache.clear_all_caches()
        host = retrieve.retrieve(who[j], queries[j])
        assert [it["chunk_id"] for it in got[j][0]] == [h["chunk_id"] for h in host]
        assert [it["combined_score"] for it in got[j][0]] == [h["combined_score"] for h in host]
    cache.clear_all_caches()
    retrieve.configure(retrieval_settings=settings_mod.RetrievalSettings())
    retrieve._DEVICE_TEXT_ENCODER = retrieve._DEVICE_IMAGE_ENCODER = retrieve._DEVICE_CROSS_ENCODER = None
    retrieve._CROSS_ENCODER = None
