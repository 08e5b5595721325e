"""GPU: the reference-signature host API (recalgorithm_b200.layers) driven exactly like the reference model_fn bodies,
with weights injected under the reference's TF variable names, against the golden vectors."""
import numpy as np
import pytest
import torch

from _util import TOL, assert_close, dev, golden

pytestmark = pytest.mark.gpu


def fresh_store():
    from recalgorithm_b200 import layers as L
    return L.set_default_store(L.VariableStore(device="cuda", seed=0))


def test_dcn_cross_part():
    from recalgorithm_b200 import layers as L
    g = golden("cross_d82_L3")
    st = fresh_store()
    st.assign({f"cross_part/wl_{i}": g["ws"][i][:, None] for i in range(3)})
    st.assign({f"cross_part/bl_{i}": g["bs"][i][:, None] for i in range(3)})
    concat_all = dev(g["x0"]).requires_grad_()
    with L.variable_scope("cross_part"):                       # DCN/dcn.py:156-160, verbatim structure
        cross_vec = concat_all
        for i in range(3):
            cross_vec = L.cross_layer(x0=concat_all, xl=cross_vec, index=i)
    assert_close(cross_vec, g["out_f64"], TOL, "per-layer cross_layer loop")
    with L.variable_scope("cross_part"):
        fused = L.cross_network(concat_all, 3)
    assert_close(fused, g["out_f64"], TOL, "fused cross_network")
    assert sorted(st.vars) == sorted([f"cross_part/{p}l_{i}" for p in "wb" for i in range(3)])
    assert st.vars["cross_part/bl_0"].shape == (82, 1)
    # gradients of both formulations agree
    g_out = torch.randn_like(cross_vec)
    ga = torch.autograd.grad(cross_vec, [concat_all, st.vars["cross_part/wl_1"]], g_out, retain_graph=True)
    gb = torch.autograd.grad(fused, [concat_all, st.vars["cross_part/wl_1"]], g_out)
    assert_close(ga[0], gb[0].double(), 2e-5, "dx0"); assert_close(ga[1], gb[1].double(), 2e-5, "dw")


def test_xdeepfm_cin_part():
    from recalgorithm_b200 import layers as L
    g = golden("cin_m8_D8_50x50x50")
    st = fresh_store()
    st.assign({f"cin_part/cin_layer_{i + 1}_filter": g[f"filter_{i + 1}"][None] for i in range(3)})
    x0 = dev(g["x0"])
    with L.variable_scope("cin_part"):                         # xDeepFM/xdeepfm.py:166-174
        xk = x0
        x_container = []
        for i, features_map_num in enumerate(["50", "50", "50"]):   # widths arrive as strings (xdeepfm.py:253)
            xk = L.cin_layer(x0, xk, features_map_num, i + 1)
            x_container.append(xk)
        p_plus = torch.cat([x.sum(dim=-1) for x in x_container], dim=-1)
    for i in range(3):
        assert_close(x_container[i], g[f"x{i + 1}_f64"], TOL, f"X^{i + 1}")
    assert_close(p_plus, g["p_plus_f64"], TOL, "p_plus")
    assert st.vars["cin_part/cin_layer_2_filter"].shape == (1, 400, 50)


@pytest.mark.parametrize("soft", [False, True])
def test_din_attention_part(soft):
    from recalgorithm_b200 import layers as L
    g = golden("din_T50")
    st = fresh_store()
    st.assign({"attention_part/f1_att/kernel": g["f1_att_kernel"], "attention_part/f1_att/bias": g["f1_att_bias"],
               "attention_part/f2_att/kernel": g["f2_att_kernel"], "attention_part/f2_att/bias": g["f2_att_bias"],
               "attention_part/f3_att/kernel": g["f3_att_kernel"], "attention_part/f3_att/bias": g["f3_att_bias"]})
    with L.variable_scope("attention_part"):                   # DIN/din.py:216-218
        out = L.din_attention(dev(g["query"]), dev(g["keys"]), dev(g["keys_length"]), is_softmax=soft)
    ref = g[f"out_softmax{int(soft)}_f64"]
    assert np.abs(out.detach().cpu().numpy() - ref).max() <= TOL * max(np.abs(ref).max(), 1e-3)
    assert len(st.vars) == 6                                    # AUTO_REUSE: a second call creates nothing new
    with L.variable_scope("attention_part"):
        L.din_attention(dev(g["query"]), dev(g["keys"]), dev(g["keys_length"]), is_softmax=soft)
    assert len(st.vars) == 6


def test_fibinet_parts():
    from recalgorithm_b200 import layers as L
    g = golden("fibinet_F8_K8")
    st = fresh_store()
    st.assign({"senet_part/senet_w1": g["senet_w1"], "senet_part/senet_w2": g["senet_w2"]})
    x = dev(g["x"])
    with L.variable_scope("senet_part"):                       # FiBiNET/fibinet.py:171-174
        senet_output = L.senet(x, embedding_dim=8, reduction_ratio=2)
    assert_close(senet_output, g["senet_f64"], TOL, "senet")
    for typ in ("all", "each", "interaction"):
        st.assign({f"bilinear_interaction_part/orginal_w_{typ}": g[f"w_{typ}"]})
        with L.variable_scope("bilinear_interaction_part"):    # FiBiNET/fibinet.py:177-181
            out = L.bilinear_interaction_layer(x, embedding_dim=8, type=typ, name="orginal")
        assert_close(out, g[f"bilinear_{typ}_f64"], TOL, typ)
    with pytest.raises(ValueError):
        L.bilinear_interaction_layer(x, 8, "nope", "x")
    with pytest.raises(AssertionError):                        # senet.py:19
        L.senet(x, embedding_dim=8, reduction_ratio=1)


def test_deepfm_fm_part_and_training_step():
    """lookup + FM2 through the autograd API: loss goes down on a toy problem (end-to-end plumbing check)."""
    from recalgorithm_b200 import autograd
    torch.manual_seed(0)
    B, F, D, rows = 256, 6, 8, 50
    tables = autograd.EmbeddingTables([rows] * F, D, device="cuda")
    ids = torch.randint(0, rows, (B, F), device="cuda")
    labels = (torch.rand((B, 1), device="cuda") < 0.3).float()
    losses = []
    for _ in range(30):
        tables.zero_grad()
        tile, fm2 = autograd.lookup_fm2(tables, ids)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(fm2, labels)
        loss.backward()
        (sl,) = tables.grad_slices
        tables.weight -= 0.5 * sl.to_dense(tables.num_rows)     # plain SGD on the densified IndexedSlices
        losses.append(loss.item())
    assert losses[-1] < losses[0] * 0.9


def test_mini_batch_aware_regularization_reaches_the_tables():
    """DIN/din.py:254-257: the L2 on looked-up activations adds lambda * e / B to the IndexedSlices of the lookup."""
    from recalgorithm_b200 import autograd, layers as L
    tables = autograd.EmbeddingTables([7, 5, 9], 8, device="cuda")
    ids = torch.tensor([[0, 4, -1], [6, 4, 8], [3, 0, 2], [0, 1, 8]], device="cuda")
    tile = autograd.lookup(tables, ids)
    B = ids.shape[0]
    reg = L.mini_batch_aware_regularization([tile.reshape(B, -1)], l2_lambda=0.2)
    assert_close(reg, 0.2 * 0.5 * float(tile.detach().double().pow(2).sum()) / B * torch.ones((), dtype=torch.float64), TOL, "value")
    reg.backward()
    assert_close(tables.grad_slices[0].values, 0.2 * tile.detach().double() / B, TOL, "lambda * e / B")


def test_c_abi_calls_are_cuda_graph_capturable():
    """Launch-bound small-batch steps (DCN at B=4096 is 4 kernels of ~10 us) can be captured into one CUDA graph: the library
    only enqueues on the given stream.  Replay with new input CONTENTS must match the eager result."""
    from recalgorithm_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(3)
    B, F, D, L, rows = 96, 5, 8, 3, 50
    table = torch.randn((rows * F, D), device="cuda", generator=gen)
    off = torch.arange(F + 1, device="cuda") * rows
    ids = torch.randint(-1, rows, (B, F), device="cuda", generator=gen)
    w, b = torch.randn((L, F * D), device="cuda", generator=gen) * 0.1, torch.randn((L, F * D), device="cuda", generator=gen) * 0.1
    g = torch.randn((B, F * D), device="cuda", generator=gen)
    outs = {}

    def step():
        tile, _ = ops.embed_fm2_fwd(table, off, ids, want_fm2=False)
        x0 = tile.view(B, F * D)
        outs["y"] = ops.cross_fwd(x0, w, b)
        dx0, _, outs["dw"], outs["db"] = ops.cross_bwd(x0, w, b, g)
        outs["rg"] = ops.embed_fm2_bwd(tile, dx0.view(B, F, D), None)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    captured = dict(outs)                                          # static output buffers of the graph
    ids.copy_(torch.randint(-1, rows, (B, F), device="cuda", generator=gen))       # new contents, same buffers
    g.copy_(torch.randn((B, F * D), device="cuda", generator=gen))
    graph.replay()
    torch.cuda.synchronize()
    got = {k: v.clone() for k, v in captured.items()}
    step()                                                         # eager, same inputs
    for k in got:
        assert torch.equal(got[k], outs[k]) or float((got[k] - outs[k]).abs().max()) <= 1e-6 * float(outs[k].abs().max()), k


def test_empty_batch_is_a_no_op_everywhere():
    """B = 0 (e.g. an empty last shard of an eval set): every op returns correctly shaped empties and ZERO parameter gradients."""
    from recalgorithm_b200 import ops
    z = lambda *s: torch.zeros(s, device="cuda")
    F, D, rows = 5, 8, 7
    table, off = torch.randn((rows * F, D), device="cuda"), torch.arange(F + 1, device="cuda") * rows
    ids = torch.zeros((0, F), dtype=torch.int64, device="cuda")
    tile, fm2 = ops.embed_fm2_fwd(table, off, ids)
    assert tile.shape == (0, F, D) and fm2.shape == (0, 1)
    assert ops.embed_fm2_bwd(tile, z(0, F, D), z(0)).shape == (0, F, D)
    t2, bi = ops.embed_bi_fwd(table, off, ids)
    assert bi.shape == (0, D) and ops.embed_bi_bwd(t2, None, z(0, D)).shape == (0, F, D)
    assert ops.first_order_fwd(torch.randn(rows * F, device="cuda"), off, ids).shape == (0, 1)
    d = F * D
    w, b = torch.randn((3, d), device="cuda"), torch.randn((3, d), device="cuda")
    assert ops.cross_fwd(z(0, d), w, b).shape == (0, d)
    dx0, _, dw, db = ops.cross_bwd(z(0, d), w, b, z(0, d))
    assert dx0.shape == (0, d) and float(dw.abs().max()) == 0 and float(db.abs().max()) == 0
    filt = torch.randn((F * F, 16), device="cuda")
    out, pooled = ops.cin_fwd(z(0, F, D), z(0, F, D), filt, want_pooled=True)
    assert out.shape == (0, 16, D) and pooled.shape == (0, 16)
    _, _, dfilt = ops.cin_bwd(z(0, F, D), z(0, F, D), filt, z(0, 16, D))
    assert float(dfilt.abs().max()) == 0
    ws = [torch.randn(s, device="cuda") for s in ((4 * D, 64), (64,), (64, 32), (32,), (32, 1), (1,))]
    lens = torch.zeros((0,), dtype=torch.int64, device="cuda")
    assert ops.din_attention_fwd(z(0, D), z(0, 6, D), lens, *ws).shape == (0, D)
    _, _, dws = ops.din_attention_bwd(z(0, D), z(0, 6, D), lens, *ws, z(0, D))
    assert all(float(x.abs().max()) == 0 for x in dws)
    w1, w2 = torch.randn((F, 4), device="cuda"), torch.randn((4, F), device="cuda")
    assert ops.senet_fwd(z(0, F, D), w1, w2).shape == (0, F, D)
    _, dw1, dw2 = ops.senet_bwd(z(0, F, D), w1, w2, z(0, F, D))
    assert float(dw1.abs().max()) == 0 and float(dw2.abs().max()) == 0
    wb = torch.randn((D, D), device="cuda")
    assert ops.bilinear_fwd(z(0, F, D), wb, "all").shape == (0, (F - 1) * (F - 2) // 2, D)
    r = torch.randn((F * (F - 1) // 2,), device="cuda")
    assert ops.fwfm_fwd(z(0, F, D), r).shape == (0, 1)
    _, dr = ops.fwfm_bwd(z(0, F, D), r, z(0))
    assert float(dr.abs().max()) == 0
    aw, ab, ah = torch.randn((D, 16), device="cuda"), torch.randn((16,), device="cuda"), torch.randn((16, 1), device="cuda")
    assert ops.afm_fwd(z(0, F, D), aw, ab, ah).shape == (0, D)
    _, daw, dab, dah = ops.afm_bwd(z(0, F, D), aw, ab, ah, z(0, D))
    assert float(daw.abs().max()) == 0 and float(dab.abs().max()) == 0 and float(dah.abs().max()) == 0
    assert ops.ffm_fwd(z(0, F, F - 1, D)).shape == (0, 1) and ops.ffm_bwd(z(0, F, F - 1, D), z(0)).shape == (0, F, F - 1, D)
    packed = torch.randn(int(ops._lib.lib().ctr_bst_param_count(D, 2, 6)), device="cuda")
    assert ops.bst_transformer_fwd(z(0, 6, D), z(0, 6, D), z(0, 6, D), lens, packed, 2, 6).shape == (0, 6, D)
    *_, dpk = ops.bst_transformer_bwd(z(0, 6, D), z(0, 6, D), z(0, 6, D), lens, packed, z(0, 6, D), 2, 6)
    assert float(dpk.abs().max()) == 0


@pytest.mark.parametrize("B", [1, 7, 256, 65536])
def test_sigmoid_ce_matches_the_reference_formula(B):
    """ctr_sigmoid_ce vs reduce_mean(sigmoid_cross_entropy_with_logits) (deepfm.py:235; stable form) in float64,
    value and gradient through autograd, with one and with two summed logits."""
    import numpy as np
    from recalgorithm_b200 import autograd
    gen = torch.Generator(device="cuda").manual_seed(B)
    a = (torch.randn((B, 1), device="cuda", generator=gen) * 3).requires_grad_()
    b = (torch.randn((B, 1), device="cuda", generator=gen) * 3).requires_grad_()
    y = (torch.rand((B, 1), device="cuda", generator=gen) < 0.3).float()
    for two in (False, True):
        a.grad = b.grad = None
        loss = autograd.sigmoid_cross_entropy_mean(a, y, logit_b=b if two else None)
        (loss * 2.0).backward()
        x = (a.detach() + (b.detach() if two else 0)).double()
        x.requires_grad_()
        ref = (torch.clamp(x, min=0) - x * y.double() + torch.log1p(torch.exp(-x.abs()))).mean()
        (ref * 2.0).backward()
        assert_close(loss.reshape(1), ref.detach().reshape(1), 1e-5, "sigmoid-CE mean")
        assert_close(a.grad, x.grad, 1e-5, "d loss / d logit")
        if two:
            assert torch.equal(a.grad, b.grad)


def _dense_table_grad(fn, table64, ids_cpu, off_cpu):
    """d loss / d table from plain float64 torch on the CPU: e = table[off[f] + id] (zero row for id < 0), loss = fn(e)."""
    t = table64.clone().requires_grad_()
    valid = (ids_cpu >= 0).unsqueeze(-1)
    e = t[(ids_cpu.clamp_min(0) + off_cpu[:-1][None, :])] * valid
    fn(e).backward()
    return t.grad


@pytest.mark.parametrize("which", ["fm2_only", "tile_only", "lin_only", "fm2_of_linear_only", "bi_only", "bi_tile_only",
                                   "cross_only", "cross_x0_only"])
def test_unused_lookup_outputs_take_no_gradient_buffer(which):
    """The multi-output lookups do not materialise a zero gradient for an output the loss never used (the backward receives None
    and hands a null pointer to the C ABI): the table gradient still equals the float64 restatement of the used branch alone."""
    from recalgorithm_b200 import autograd
    torch.manual_seed(7)
    B, F, D, rows, L = 37, 6, 8, 11, 2
    tables = autograd.EmbeddingTables([rows] * F, D, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    ids = torch.randint(-1, rows, (B, F), device="cuda")
    t64, ids_c, off_c = tables.weight.double().cpu(), ids.cpu(), tables.field_row_offset.cpu()
    gt = torch.randn((B, F, D), dtype=torch.float64)
    wl = torch.randn((F * D,), dtype=torch.float64) * 0.1
    cw, cb = torch.randn((L, F * D), dtype=torch.float64) * 0.05, torch.randn((L, F * D), dtype=torch.float64) * 0.05
    gx = torch.randn((B, F * D), dtype=torch.float64)

    def fm2(e):
        return 0.5 * ((e.sum(1) ** 2).sum(1) - (e ** 2).sum((1, 2)))

    def bi(e):
        return 0.5 * (e.sum(1) ** 2 - (e ** 2).sum(1))

    def cross(e):
        x0 = e.reshape(B, -1); xl = x0
        for l in range(L):
            xl = x0 * (xl @ cw[l])[:, None] + cb[l] + xl
        return xl

    f32 = lambda x: x.float().cuda()                                          # noqa: E731
    if which == "fm2_only":
        _, out = autograd.lookup_fm2(tables, ids); loss = out.sum(); ref = lambda e: fm2(e).sum()                     # noqa: E731
    elif which == "tile_only":
        out, _ = autograd.lookup_fm2(tables, ids); loss = (out * f32(gt)).sum(); ref = lambda e: (e * gt).sum()       # noqa: E731
    elif which in ("lin_only", "fm2_of_linear_only"):
        w = f32(wl).requires_grad_()
        o_fm2, o_lin = autograd.lookup_fm2_linear(tables, ids, w)
        if which == "lin_only":
            loss = o_lin.sum(); ref = lambda e: (e.reshape(B, -1) @ wl).sum()                                          # noqa: E731
        else:
            loss = o_fm2.sum(); ref = lambda e: fm2(e).sum()                                                          # noqa: E731
    elif which == "bi_only":
        _, out = autograd.lookup_bi(tables, ids); loss = out.sum(); ref = lambda e: bi(e).sum()                       # noqa: E731
    elif which == "bi_tile_only":
        out, _ = autograd.lookup_bi(tables, ids); loss = (out * f32(gt)).sum(); ref = lambda e: (e * gt).sum()        # noqa: E731
    else:
        w, b = f32(cw).requires_grad_(), f32(cb).requires_grad_()
        xl, x0 = autograd.lookup_cross(tables, ids, w, b)
        if which == "cross_only":
            loss = (xl * f32(gx)).sum(); ref = lambda e: (cross(e) * gx).sum()                                        # noqa: E731
        else:
            loss = (x0 * f32(gx)).sum(); ref = lambda e: (e.reshape(B, -1) * gx).sum()                                # noqa: E731
    loss.backward()
    (sl,) = tables.grad_slices
    assert_close(sl.to_dense(tables.num_rows), _dense_table_grad(ref, t64, ids_c, off_c), TOL, which)
    if which == "lin_only":
        e = t64[(ids_c.clamp_min(0) + off_c[:-1][None, :])] * (ids_c >= 0).unsqueeze(-1)
        assert_close(w.grad, e.reshape(B, -1).sum(0), TOL, "d_wlin")
    if which == "fm2_of_linear_only":
        assert float(w.grad.abs().max()) == 0.0                               # d_lin = None -> d_wlin is exactly zero
    if which == "cross_x0_only":
        assert w.grad is None and b.grad is None                              # the cross weights were not on the used branch
