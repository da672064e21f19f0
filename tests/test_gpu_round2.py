"""Round-2 behaviour of the scan library: grouped varlen launch (K6), host-buffer mailbox path, precision policy,
argument checks added after the round-1 review."""
import importlib
import threading

import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu
PKG = "multimodal-rag-for-image-text-search_b200"


@pytest.fixture(scope="module")
def mmr():
    return importlib.import_module(PKG)


@pytest.fixture(scope="module")
def table(mmr):
    rows = util.unit_rows(120_000, 512, seed=301)
    seg = np.array([0, 7, 7, 30_000, 30_001, 90_000, 120_000], dtype=np.int64)
    ix = mmr.ResidentIndex.from_f32(rows, seg_offsets=seg, dtype="bf16")
    yield rows, seg, ix
    ix.close()


def test_grouped_varlen_launch_equals_single_queries(mmr, table):
    """K6: queries of one tenant share a pass over its rows (groups of 4); every result is bit-identical to the same
    query searched alone (same fp32 arithmetic), whatever it is batched with."""
    rows, seg, ix = table
    tenants = [2, 4, 2, 2, 0, 5, 2, 4, 2, 2, 1, 3, 2, 4, -1, 2, 2, 5]      # tenant 2 nine times: groups of 4 + 4 + 1
    q = torch.from_numpy(util.queries(len(tenants), 512, seed=302)).cuda()
    for k in (10, 50):
        s, r = ix.search(q, k, tenants)
        assert mmr._native.lib().mmr_last_kernel() == 3
        for j, t in enumerate(tenants):
            s1, r1 = ix.search(q[j:j + 1], k, [t])
            assert mmr._native.lib().mmr_last_kernel() == 1
            assert torch.equal(r[j], r1[0]) and torch.equal(s[j], s1[0]), (k, j, t)
        lo, hi = int(seg[2]), int(seg[3])
        util.check_topk(s[0].cpu().numpy(), r[0].cpu().numpy(), util.oracle_scores(rows[lo:hi], q[0].cpu().numpy()), k,
                        util.TOL_BF16, lo=lo, hi=hi, what="grouped varlen")


def test_precision_policy_f32_keeps_batches_bit_identical(mmr, table):
    """the same (tenant, query) must not score differently depending on its batch.  With the serving policy
    (fp32 queries) a same-tenant batch of 9 runs as K1 passes and equals the single-query results bit for bit; the
    throughput policy ("auto") sends it to the tensor cores (16-bit queries, scores within 1e-3)."""
    rows, seg, ix = table
    q = torch.from_numpy(util.queries(9, 512, seed=303)).cuda()
    singles = [ix.search(q[j:j + 1], 10, [4]) for j in range(9)]
    ix.set_query_precision("f32")
    s, r = ix.search(q, 10, [4] * 9)
    assert mmr._native.lib().mmr_last_kernel() == 1
    for j in range(9):
        assert torch.equal(r[j], singles[j][1][0]) and torch.equal(s[j], singles[j][0][0])
    ix.set_query_precision("auto")
    s2, r2 = ix.search(q, 10, [4] * 9)
    assert mmr._native.lib().mmr_last_kernel() == 2
    assert (s2 - s).abs().max().item() < util.TOL_BF16


def test_rescore_policy_is_bit_identical_to_single_queries_on_the_tensor_cores(mmr, table):
    """The serving default: tensor-core candidates + fp32 re-scoring in K1's arithmetic + an exactness proof (rerun on K1
    when it fails).  Same-tenant batches then run on the tensor cores AND equal the single-query results bit for bit."""
    rows, seg, ix = table
    lib = mmr._native.lib()
    ix.set_query_precision("rescore")
    try:
        for tenant, b, k in ((4, 9, 10), (4, 130, 12), (2, 40, 50), (-1, 300, 10), (0, 5, 10), (3, 4, 3), (5, 7, 64)):
            q = torch.from_numpy(util.queries(b, 512, seed=400 + b)).cuda()
            q[1] = q[0] * 3.0                                    # not unit norm: re-normalised on the device
            s, r = ix.search(q, k, [tenant] * b)
            assert lib.mmr_last_kernel() == 2, (tenant, b, k)
            ix.set_query_precision("f32")
            for j in range(0, b, max(1, b // 12)):
                s1, r1 = ix.search(q[j:j + 1], k, [tenant])
                assert torch.equal(r[j], r1[0]) and torch.equal(s[j], s1[0]), (tenant, b, k, j)
            ix.set_query_precision("rescore")
        # duplicates: exact ties between a nominated and a non-nominated row cannot flip the order
        q = torch.from_numpy(np.repeat(rows[40_000:40_001], 6, axis=0)).cuda()
        s, r = ix.search(q, 10, [4] * 6)
        ix.set_query_precision("f32")
        s1, r1 = ix.search(q[:1], 10, [4])
        assert torch.equal(r[3], r1[0]) and torch.equal(s[3], s1[0]) and int(r[3, 0]) == 40_000
        print("rescore reruns so far:", lib.mmr_rescore_reruns())
    finally:
        ix.set_query_precision("auto")


def test_store_same_tenant_batch_equals_per_request(mmr):
    """Through the drop-in: 6 requests of one tenant in one micro-batch == the 6 requests served alone, dict for dict."""
    rng = np.random.default_rng(304)
    n = 20_000
    emb = util.unit_rows(n, 512, seed=305)
    store = mmr.B200Store()
    store.load_arrow("image_collection", mmr.make_arrow_table([f"i{i}" for i in range(n)], ["u" if i % 4 else "v" for i in range(n)],
                                                              ["d"] * n, ["image"] * n, emb, ["{}"] * n))
    qs = rng.standard_normal((6, 512)).astype(np.float32)
    batch = store.search_image_batch(["u"] * 6, qs, 12)
    for j in range(6):
        assert batch[j] == store.search_image("u", qs[j].tolist(), 12)
    mixed = store.search_image_batch(["u", "v", "u", "u", "nobody", "v"], qs, 12)
    for j, u in enumerate(["u", "v", "u", "u", "nobody", "v"]):
        assert mixed[j] == store.search_image(u, qs[j].tolist(), 12)


def test_overlapping_ranges_are_rejected(mmr, table):
    _, _, ix = table
    q = torch.from_numpy(util.queries(2, 512, seed=306)).cuda()
    ix.search_ranges(q, 10, [[(0, 100), (100, 200)], [(5, 9)]])            # touching is fine
    with pytest.raises(mmr.NativeError, match="overlap"):
        ix.search_ranges(q, 10, [[(0, 100), (50, 200)], [(5, 9)]])


def test_zero_query_in_a_tensor_core_batch_returns_k_rows(mmr, table):
    """the probe floor of an all-zero query is +0.0; the float below it must not be -0.0 (which would reject
    every row scoring exactly 0).  K2 must return the first k rows at score 0 like K1 does."""
    rows, seg, ix = table
    q = util.queries(8, 512, seed=307)
    q[3] = 0.0
    s, r = ix.search(torch.from_numpy(q).cuda(), 10)
    assert mmr._native.lib().mmr_last_kernel() == 2
    assert r[3].cpu().tolist() == list(range(10)) and (s[3] == 0).all()
    s1, r1 = ix.search(torch.from_numpy(q[3:4]).cuda(), 10)
    assert torch.equal(r1[0], r[3])


def test_fragmented_tenants_on_a_large_table_fit_the_workspace(mmr):
    """40 queries x 8 row ranges each on a ~0.7M-row scan used to overflow the varlen item table."""
    rows = util.unit_rows(700_000, 384, seed=308)
    ix = mmr.ResidentIndex.from_f32(rows, dtype="bf16")
    rng = np.random.default_rng(309)
    ranges = []
    for b in range(40):
        cuts = np.sort(rng.choice(700_000, size=16, replace=False))
        ranges.append([(int(cuts[2 * i]), int(cuts[2 * i + 1])) for i in range(8)])
    q = util.queries(40, 384, seed=310)
    s, r = ix.search_ranges(torch.from_numpy(q).cuda(), 10, ranges)
    s, r = s.cpu().numpy(), r.cpu().numpy()
    for b in (0, 13, 39):
        mask = np.zeros(700_000, bool)
        for lo, hi in ranges[b]:
            mask[lo:hi] = True
        full = util.oracle_scores(rows, q[b])
        full[~mask] = -np.inf
        util.check_topk(s[b], r[b], full, 10, util.TOL_BF16, what=f"fragmented q{b}")
    ix.close()


def test_host_calls_mailbox_and_reentrancy(mmr, table):
    """mmr_search_host: B <= 2 on one range -> one launch, query in the kernel parameters, result + flag in the mapped
    mailbox; other shapes stage the queries and synchronise.  All equal the device-buffer path; concurrent callers on one
    index take turns (ADVICE r1 / VERDICT weak #9)."""
    rows, seg, ix = table
    lib = mmr._native.lib()
    q = util.queries(6, 512, seed=311)
    qd = torch.from_numpy(q).cuda()
    for mailbox, inline in (("0", "1"), ("1", "1"), ("1", "0")):   # completion by stream sync / by the kernel-written flag
        mmr._native.set_option("MMR_MAILBOX", mailbox)
        mmr._native.set_option("MMR_INLINE_QUERY", inline)
        for b, segs in ((1, [4]), (2, [4, 4]), (2, None), (5, [4] * 5), (6, [0, 2, 4, 5, 2, -1]), (1, [1])):
            n0 = lib.mmr_launch_count()
            hs, hr = ix.search_host(q[:b], 10, segs)
            launches = lib.mmr_launch_count() - n0
            ds, dr = ix.search(qd[:b], 10, segs)
            assert (hr == dr.cpu().numpy()).all() and (hs == ds.cpu().numpy()).all(), (b, segs, mailbox, inline)
            if b <= 2:
                assert launches == 1, "the single-request path is one kernel launch"
    mmr._native.set_option("MMR_MAILBOX", None)
    mmr._native.set_option("MMR_INLINE_QUERY", None)
    errors = []

    def worker(j):
        try:
            for rep in range(40):
                hs, hr = ix.search_host(q[j:j + 1], 10, [4])
                ds, dr = ix.search(qd[j:j + 1], 10, [4])       # separate workspace use is serialised by this thread's stream order
                assert (hr == dr.cpu().numpy()).all()
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)

    # host calls from several threads (the device-buffer calls inside share self._ws, so keep those on one thread)
    def host_only(j):
        try:
            want = ix.search_host(q[j:j + 1], 10, [4])[1]
            for rep in range(60):
                assert (ix.search_host(q[j:j + 1], 10, [4])[1] == want).all()
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=host_only, args=(j,)) for j in range(4)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors
    worker(0)
    assert not errors, errors
