#!/usr/bin/env python
"""Per-layer timings of the non-headline BASELINE configs (2: DCN cross, 3: xDeepFM CIN, 4: DIN attention) and FiBiNET.

Each kernel is timed alone with CUDA events; an L2 flush (a 256 MB write) runs between timed iterations because these
working sets fit the 126 MB L2.  Prints one JSON line per measurement; meant to be redirected into profiles/.

    python tools/bench_layers.py [--iters 20] [--only cin,dcn,din,fibinet]
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from recalgorithm_b200 import ops  # noqa: E402


def peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return d["hbm_gbs"], d["bf16_tflops"], "measured"
    except Exception:
        return 6650.0, 1590.0, "fallback"


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        flush.zero_()                                   # 256 MB write: evicts the 126 MB L2
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    return statistics.median(times), min(times)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default="dcn,cin,din,fibinet")
    ap.add_argument("--sweep", action="store_true", help="bilinear: also time every (columns per lane, tile) variant of the tournament kernels")
    args = ap.parse_args()
    only = set(args.only.split(","))
    torch.cuda.set_device(0)
    gen = torch.Generator(device="cuda").manual_seed(1234)
    rn = lambda *s, std=1.0: torch.randn(s, device="cuda", generator=gen) * std
    flush = torch.empty(64 * 1024 * 1024, device="cuda")
    hbm, tf, src = peaks()

    def emit(name, cfg, med, best, bytes_=None, flops=None, note=""):
        line = {"kernel": name, "config": cfg, "ms_median": med, "ms_best": best, "l2": "flushed between iterations", "note": note}
        if bytes_ is not None:
            line["algorithmic_GBps"] = bytes_ / (med * 1e-3) / 1e9
            line["frac_of_hbm_peak"] = line["algorithmic_GBps"] / hbm
        if flops is not None:
            line["TFLOPs"] = flops / (med * 1e-3) / 1e12
            line["frac_of_bf16_peak"] = line["TFLOPs"] / tf
        line["peak_source"] = src
        print(json.dumps(line), flush=True)

    if "dcn" in only:   # BASELINE config 2: DCN 3 cross layers, 30 fields x 16 = 480, batch 4096
        B, d, L = 4096, 480, 3
        x0, w, b, g = rn(B, d), rn(L, d, std=0.05), rn(L, d, std=0.05), rn(B, d)
        cfg = {"B": B, "d": d, "L": L}
        m, bst = timeit(lambda: ops.cross_fwd(x0, w, b), args.iters, flush)
        emit("cross_fwd", cfg, m, bst, bytes_=B * 2 * d * 4, note="2*d*4 B/sample")
        m, bst = timeit(lambda: ops.cross_bwd(x0, w, b, g), args.iters, flush)
        emit("cross_bwd", cfg, m, bst, bytes_=B * 3 * d * 4, note="3*d*4 B/sample")
        for B2 in (65536,):
            x0b, gb = rn(B2, d), rn(B2, d)
            m, bst = timeit(lambda: ops.cross_fwd(x0b, w, b), args.iters, flush)
            emit("cross_fwd", {"B": B2, "d": d, "L": L}, m, bst, bytes_=B2 * 2 * d * 4)
            m, bst = timeit(lambda: ops.cross_bwd(x0b, w, b, gb), args.iters, flush)
            emit("cross_bwd", {"B": B2, "d": d, "L": L}, m, bst, bytes_=B2 * 3 * d * 4)

    if "cin" in only:   # BASELINE config 3: xDeepFM CIN [128,128], 30 fields, D=16, batch 8192
        B, mm, D, H = 8192, 30, 16, 128
        x0 = rn(B, mm, D, std=0.25)
        w1, w2 = rn(mm * mm, H, std=0.05), rn(H * mm, H, std=0.05)
        x1 = ops.cin_fwd(x0, x0, w1)
        g = rn(B, H, D)
        for prec in (0, 1):
            tag = "3xTF32" if prec == 0 else "1xTF32"
            m, bst = timeit(lambda: ops.cin_fwd(x0, x0, w1, want_pooled=True, precision=prec), args.iters, flush)
            emit(f"cin_fwd_layer1_{tag}", {"B": B, "m": mm, "hk": mm, "D": D, "H": H}, m, bst, flops=2.0 * B * D * mm * mm * H)
            m, bst = timeit(lambda: ops.cin_fwd(x0, x1, w2, want_pooled=True, precision=prec), args.iters, flush)
            emit(f"cin_fwd_layer2_{tag}", {"B": B, "m": mm, "hk": H, "D": D, "H": H}, m, bst, flops=2.0 * B * D * H * mm * H)
        for pair in (1, 0):
            ops._lib.lib().ctr_cin_bwd_set_dx_pair(pair)
            tag = "dx=cta_group::2 pairs" if pair else "dx=single CTA + multicast"
            m, bst = timeit(lambda: ops.cin_bwd(x0, x0, w1, g), max(3, args.iters // 4), flush)
            emit("cin_bwd_layer1", {"B": B, "m": mm, "hk": mm, "D": D, "H": H, "variant": tag}, m, bst, flops=4.0 * B * D * mm * mm * H)
            m, bst = timeit(lambda: ops.cin_bwd(x0, x1, w2, g), max(3, args.iters // 4), flush)
            emit("cin_bwd_layer2", {"B": B, "m": mm, "hk": H, "D": D, "H": H, "variant": tag}, m, bst, flops=4.0 * B * D * H * mm * H)
        ops._lib.lib().ctr_cin_bwd_set_dx_pair(1)

    if "din" in only:   # BASELINE config 4: DIN attention, seq_len 50, H=16, batch 4096
        B, T, H = 4096, 50, 16
        q, k = rn(B, H, std=0.25), rn(B, T, H, std=0.25)
        lens = torch.randint(0, T + 1, (B,), device="cuda", generator=gen)
        ws = [rn(4 * H, 64, std=0.2), rn(64, std=0.1), rn(64, 32, std=0.2), rn(32, std=0.1), rn(32, 1, std=0.3), rn(1, std=0.1)]
        g = rn(B, H)
        flops_full = B * T * 2.0 * (4 * H * 64 + 64 * 32 + 32)
        for soft in (False, True):
            m, bst = timeit(lambda: ops.din_attention_fwd(q, k, lens, *ws, is_softmax=soft), args.iters, flush)
            emit(f"din_fwd_softmax{int(soft)}", {"B": B, "T": T, "H": H, "lengths": "uniform{0..50}"}, m, bst,
                 bytes_=B * (T * H * 4 + 8 + 2 * H * 4), flops=flops_full,
                 note="FLOPs counted as the reference does them (all T positions, unfolded layer 1)")
            m, bst = timeit(lambda: ops.din_attention_bwd(q, k, lens, *ws, g, is_softmax=soft), args.iters, flush)
            emit(f"din_bwd_softmax{int(soft)}", {"B": B, "T": T, "H": H}, m, bst, flops=2 * flops_full)

    if "fibinet" in only:   # FiBiNET F=30, K=16, r=8 (no BASELINE config; reference defaults)
        B, F, K, r = 4096, 30, 16, 8
        x, w1, w2 = rn(B, F, K, std=0.25), rn(F, r, std=0.3), rn(r, F, std=0.3)
        g = rn(B, F, K)
        m, bst = timeit(lambda: ops.senet_fwd(x, w1, w2), args.iters, flush)
        emit("senet_fwd", {"B": B, "F": F, "K": K, "r": r}, m, bst, bytes_=B * 2 * F * K * 4)
        m, bst = timeit(lambda: ops.senet_bwd(x, w1, w2, g), args.iters, flush)
        emit("senet_bwd", {"B": B, "F": F, "K": K, "r": r}, m, bst, bytes_=B * 3 * F * K * 4)
        P = (F - 1) * (F - 2) // 2
        gp = rn(B, P, K)
        for typ in ("all", "each", "interaction"):
            w = rn(*ops.bilinear_w_shape(F, K, typ), std=0.2)
            variants = [(4, "default (staged all/each, tournament interaction)"), (7, "tournament")]
            if args.sweep:
                variants += [(7 | (kt << 10) | (tile << 4), f"tournament kt={kt} tile={tile}") for kt in (1, 2, 4) for tile in (4, 8, 16)]
            variants += [(8, "round-1 CTA-per-sample kernels")]
            for mask, impl in variants:
                prev = ops.bilinear_set_tournament(mask)
                m, bst = timeit(lambda: ops.bilinear_fwd(x, w, typ), args.iters, flush)
                emit(f"bilinear_fwd_{typ}", {"B": B, "F": F, "K": K, "impl": impl}, m, bst, bytes_=B * (F * K + P * K) * 4)
                m, bst = timeit(lambda: ops.bilinear_bwd(x, w, typ, gp), max(3, args.iters // 4), flush)
                emit(f"bilinear_bwd_{typ}", {"B": B, "F": F, "K": K, "impl": impl}, m, bst, bytes_=B * (2 * F * K + P * K) * 4)
                ops.bilinear_set_tournament(prev)

    if "pairwise" in only:   # SURVEY 8f.4 siblings (FwFM at the config-5 tile shape; AFM at the reference's flag defaults and at F=30)
        B, F, K = 65536, 40, 32
        x, r, g1 = rn(B, F, K, std=0.2), rn(F * (F - 1) // 2, std=0.3), rn(B)
        m_, bst = timeit(lambda: ops.fwfm_fwd(x, r), args.iters, flush)
        emit("fwfm_fwd", {"B": B, "F": F, "K": K}, m_, bst, bytes_=B * F * K * 4, flops=2.0 * B * K * F * (F - 1) / 2)
        m_, bst = timeit(lambda: ops.fwfm_bwd(x, r, g1), args.iters, flush)
        emit("fwfm_bwd", {"B": B, "F": F, "K": K}, m_, bst, bytes_=2 * B * F * K * 4, flops=2.0 * B * K * (F * F + F * (F - 1) / 2))
        for (B, F, K, T) in ((65536, 7, 8, 128), (8192, 30, 16, 8)):
            x, w, b, h, gk = rn(B, F, K, std=0.5), rn(K, T, std=0.3), rn(T, std=0.1), rn(T, std=0.3), rn(B, K)
            P = F * (F - 1) // 2
            m_, bst = timeit(lambda: ops.afm_fwd(x, w, b, h), args.iters, flush)
            emit("afm_fwd", {"B": B, "F": F, "K": K, "T": T}, m_, bst, bytes_=B * (F * K + K) * 4, flops=2.0 * B * P * (K * T + T + K))
            m_, bst = timeit(lambda: ops.afm_bwd(x, w, b, h, gk), args.iters, flush)
            emit("afm_bwd", {"B": B, "F": F, "K": K, "T": T}, m_, bst, bytes_=B * (2 * F * K + K) * 4, flops=2.0 * B * P * (4 * K * T + 2 * T + 4 * K))

    if "bst" in only:   # BST defaults: 50-step history + target = 51 positions, d = 8, 3 heads (BST/bst.py:45-47), batch 4096
        for (B, T, d, H) in ((4096, 51, 8, 3), (4096, 51, 16, 3)):
            x, g = rn(B, T, d), rn(B, T, d)
            klen = torch.randint(1, T + 1, (B,), device="cuda", generator=gen)
            packed = rn(int(ops._lib.lib().ctr_bst_param_count(d, H, T)), std=0.3)
            flops = 2.0 * B * (3 * H * T * d * d + 2 * H * T * T * d + H * T * d * d + T * d * d)
            m_, bst = timeit(lambda: ops.bst_transformer_fwd(x, x, x, klen, packed, H, T), args.iters, flush)
            emit("bst_transformer_fwd", {"B": B, "T": T, "d": d, "heads": H}, m_, bst, bytes_=2 * B * T * d * 4, flops=flops)
            m_, bst = timeit(lambda: ops.bst_transformer_bwd(x, x, x, klen, packed, g, H, T), args.iters, flush)
            emit("bst_transformer_bwd", {"B": B, "T": T, "d": d, "heads": H}, m_, bst, bytes_=5 * B * T * d * 4, flops=3 * flops)

    if "adam" in only:   # SURVEY 8f.3 at BASELINE config 5: the table update that follows the hot path
        from recalgorithm_b200 import autograd, optim
        rows = int(os.environ.get("CTR_BENCH_ROWS", 2_500_000))
        B, F, D = 65536, 40, 32
        tables = autograd.EmbeddingTables([rows] * F, D, device="cuda")
        ids = [torch.randint(0, rows, (B, F), device="cuda", generator=gen) for _ in range(4)]
        vals = rn(B, F, D)
        for lazy in (True, False):
            opt = optim.TableAdam(tables, lr=1e-3, lazy=lazy)
            k = [0]

            def step():
                k[0] += 1
                tables.grad_slices.append(autograd.IndexedSlices(vals.clone(), ids[k[0] % 4], tables.field_row_offset))
                opt.step()
            m_, bst = timeit(step, max(5, args.iters // 2), flush)
            n = B * F
            # algorithmic bytes: ids + values read; (m, v, var) read+written for the referenced rows; dense variant: 6 table streams
            by = n * 8 + n * D * 4 + 6 * n * D * 4 if lazy else 6 * tables.num_rows * D * 4 + n * 8 + n * D * 4
            emit("table_adam_lazy" if lazy else "table_adam_dense", {"B": B, "F": F, "D": D, "rows_per_field": rows}, m_, bst,
                 bytes_=by, note="includes a 336 MB clone of the gradient values per step (the step consumes them)")
            del opt
        # backward + LazyAdam: unfused pair (ctr_embed_fm2_bwd writes row_grads, ctr_adam_indexed_slices re-reads them) vs the
        # fused kernel (ctr_embed_fm2_bwd_adam)
        tile, _ = ops.embed_fm2_fwd(tables.weight, tables.field_row_offset, ids[0])
        d_tile, d_fm2 = rn(B, F, D, std=0.01), rn(B, std=0.01)
        opt_u = optim.TableAdam(tables, lr=1e-3, lazy=True)
        rg = torch.empty_like(tile)

        def unfused():
            k[0] += 1
            ops.embed_fm2_bwd(tile, d_tile, d_fm2, row_grads=rg)
            tables.grad_slices.append(autograd.IndexedSlices(rg, ids[k[0] % 4], tables.field_row_offset))
            opt_u.step()
        m_, bst = timeit(unfused, max(5, args.iters // 2), flush)
        n = B * F
        by_u = n * 8 + 3 * n * D * 4 + n * D * 4 + 6 * n * D * 4
        emit("bwd_plus_lazy_adam_unfused", {"B": B, "F": F, "D": D, "rows_per_field": rows}, m_, bst, bytes_=by_u,
             note="embed_fm2_bwd (read tile, d_tile; write row_grads) + claim/merge/update (read row_grads; RMW m, v, var)")
        del opt_u
        opt_f = optim.TableAdam(tables, lr=1e-3, lazy=True, fused_backward=True)

        def fused():
            k[0] += 1
            opt_f.apply_fused(tile, d_tile, d_fm2, ids[k[0] % 4])
            opt_f.step()
        m_, bst = timeit(fused, max(5, args.iters // 2), flush)
        by_f = n * 8 + 2 * n * D * 4 + 6 * n * D * 4
        emit("bwd_plus_lazy_adam_fused", {"B": B, "F": F, "D": D, "rows_per_field": rows}, m_, bst, bytes_=by_f,
             note="ctr_embed_fm2_bwd_adam: read tile, d_tile; RMW m, v, var; row_grads only for rows with duplicates")


if __name__ == "__main__" and "--configs" not in sys.argv:
    main()


def config_level():
    """BASELINE configs 2-4 as lookup + interaction chains (hot path only: no dense tail), forward+backward, samples/s.
    Timed as one CUDA-event region per step, L2 flushed between steps; tables 1 M rows per field."""
    import statistics
    torch.cuda.set_device(0)
    gen = torch.Generator(device="cuda").manual_seed(1234)
    flush = torch.empty(64 * 1024 * 1024, device="cuda")
    rn = lambda *s, std=1.0: torch.randn(s, device="cuda", generator=gen) * std

    def run(name, cfg, B, step, iters=15):
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); step(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = statistics.median(ts)
        print(json.dumps({"config_level": name, "config": cfg, "ms_per_step": ms, "samples_per_sec": B / (ms * 1e-3),
                          "what": "lookup fwd + interaction fwd + interaction bwd + lookup bwd (IndexedSlices); no dense tail",
                          "l2": "flushed between steps"}), flush=True)

    rows = 1_000_000
    # ---- config 2: DCN, 30 fields x 16, L = 3, B = 4096
    B, F, D, L = 4096, 30, 16, 3
    table = rn(rows * F, D, std=D ** -0.5); off = torch.arange(F + 1, device="cuda") * rows
    ids = torch.randint(0, rows, (B, F), device="cuda", generator=gen)
    w, bb = rn(L, F * D, std=0.05), rn(L, F * D, std=0.05); g = rn(B, F * D)

    def dcn():
        tile, _ = ops.embed_fm2_fwd(table, off, ids, want_fm2=False)
        x0 = tile.view(B, F * D)
        ops.cross_fwd(x0, w, bb)
        dx0, _, _, _ = ops.cross_bwd(x0, w, bb, g)
        ops.embed_fm2_bwd(tile, dx0.view(B, F, D), None)
    run("dcn_cfg2", {"B": B, "F": F, "D": D, "L": L, "rows_per_field": rows}, B, dcn)

    def graphed(fn):
        """The same C-ABI calls captured once into a CUDA graph (they only enqueue work on the given stream and never
        allocate or synchronise, so they are capturable as they are); a step is then ONE graph launch."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            fn()
        return graph.replay
    run("dcn_cfg2_cuda_graph", {"B": B, "F": F, "D": D, "L": L, "rows_per_field": rows, "launch": "one CUDA graph replay per step"}, B,
        graphed(dcn))

    # ---- config 3: xDeepFM CIN [128,128], 30 fields x 16, B = 8192
    B, F, D, H = 8192, 30, 16, 128
    ids3 = torch.randint(0, rows, (B, F), device="cuda", generator=gen)
    w1, w2 = rn(F * F, H, std=0.05), rn(H * F, H, std=0.05)
    gp = rn(B, 2 * H)

    def xdeepfm():
        x0, _ = ops.embed_fm2_fwd(table, off, ids3, want_fm2=False)
        x1, p1 = ops.cin_fwd(x0, x0, w1, want_pooled=True)
        x2, p2 = ops.cin_fwd(x0, x1, w2, want_pooled=True)
        g2 = gp[:, H:].unsqueeze(-1).expand(B, H, D).contiguous()          # d(pooled)/d(out) broadcast over D
        dx0b, dx1, _ = ops.cin_bwd(x0, x1, w2, g2)
        g1 = dx1 + gp[:, :H].unsqueeze(-1)
        dx0a, dxk, _ = ops.cin_bwd(x0, x0, w1, g1.contiguous())
        ops.embed_fm2_bwd(x0, (dx0a + dxk + dx0b).contiguous(), None)
    run("xdeepfm_cfg3", {"B": B, "m": F, "D": D, "cin": [H, H], "rows_per_field": rows}, B, xdeepfm)

    # ---- config 4: DIN attention, history T = 50 (one lookup per step), H = 16, B = 4096
    B, T, Hd = 4096, 50, 16
    tab4 = rn(rows, Hd, std=0.25); off4 = torch.tensor([0, rows], device="cuda")
    lens = torch.randint(0, T + 1, (B,), device="cuda", generator=gen)
    hist = torch.randint(0, rows, (B * T, 1), device="cuda", generator=gen)
    hist[(torch.arange(T, device="cuda")[None, :] >= lens[:, None]).reshape(-1)] = -1     # padding -> zero vectors
    tgt = torch.randint(0, rows, (B, 1), device="cuda", generator=gen)
    ws = [rn(4 * Hd, 64, std=0.2), rn(64, std=0.1), rn(64, 32, std=0.2), rn(32, std=0.1), rn(32, 1, std=0.3), rn(1, std=0.1)]
    go = rn(B, Hd)

    def din():
        keys, _ = ops.embed_fm2_fwd(tab4, off4, hist, want_fm2=False)
        q, _ = ops.embed_fm2_fwd(tab4, off4, tgt, want_fm2=False)
        k3, q2 = keys.view(B, T, Hd), q.view(B, Hd)
        out, att = ops.din_attention_fwd(q2, k3, lens, *ws, want_weights=True)
        dq, dk, _ = ops.din_attention_bwd(q2, k3, lens, *ws, go, att_w=att)
        ops.embed_fm2_bwd(keys, dk.view(B * T, 1, Hd), None)
        ops.embed_fm2_bwd(q, dq.view(B, 1, Hd), None)
    run("din_cfg4", {"B": B, "T": T, "H": Hd, "rows": rows, "lengths": "uniform{0..50}"}, B, din)


if __name__ == "__main__" and "--configs" in sys.argv:
    config_level()
