"""Multi-GPU worker (launched by torchrun, one rank per GPU): peer-pull lookup and fused gradient push vs the
pure-torch reference exchange.  Prints 'SHARDED_OK rank=..' on success."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from oracle import sharded_ref as R                      # checker only
    from recalgorithm_b200 import ops, sharded as S
    dev = torch.device("cuda", local)
    for ci, (B, F, D, rows_each) in enumerate(((257, 40, 32, 5000), (64, 6, 8, 300), (1000, 33, 16, 1))):
        g = torch.Generator(device=dev).manual_seed(11)
        rows = [rows_each + 3 * f for f in range(F)]
        # both shard allocators: torch symmetric memory and the library's own VMM allocation (fd passing)
        t = S.ShardedEmbeddingTables(rows, D, batch_per_rank=B, device=dev, init=None, slack=3.0,
                                     shard_backend="vmm" if ci % 2 else "symm")
        full = torch.randn((t.num_rows, D), device=dev, generator=g)                # identical on all ranks
        t.weight.copy_(S.full_to_shard(full, rank, world))
        dist.barrier()
        gi = torch.Generator(device=dev).manual_seed(1000 + rank)
        rt = torch.tensor(rows, device=dev)
        ids = (torch.rand((B, F), device=dev, generator=gi) * (rt[None, :] + 3)).long() - 1   # includes -1 and >= rows
        tile, fm2 = t.lookup_fm2(ids)
        valid = (ids >= 0) & (ids < rt[None, :])
        want = full[(ids + t.field_row_offset[:-1][None, :]).clamp(0, t.num_rows - 1)] * valid[..., None]
        assert torch.equal(tile, want), "peer-pull lookup must be an exact copy"
        e = want.double()
        ref = 0.5 * (e.sum(1).pow(2) - e.pow(2).sum(1)).sum(1, keepdim=True)
        assert (fm2.double() - ref).abs().max() <= 1e-5 * ref.abs().max().clamp_min(1e-30)
        d_tile = torch.randn((B, F, D), device=dev, generator=gi)
        d_fm2 = torch.randn((B,), device=dev, generator=gi)
        row_grads = ops.embed_fm2_bwd(tile, d_tile, d_fm2)
        refd = R.exchange_reference(t.local_rows, t.field_row_offset, ids, row_grads)
        # (1) plain push of existing row gradients
        t.push_grads(ids, row_grads)
        got = t.received_to_dense()
        err = (got.double() - refd).abs().max() / refd.abs().max().clamp_min(1e-30)
        assert err <= 1e-5, f"pushed gradients differ: {err}"
        assert int(t.recv_counts.sum()) > 0
        dist.barrier()
        # (2) backward fused with the push: row_grads never materialised
        t.recv_vals.zero_()
        torch.cuda.synchronize(); dist.barrier()
        plan = t.plan(ids)
        t.bwd_push(tile, d_tile, d_fm2, plan)
        t.finish_push()
        got = t.received_to_dense()
        err = (got.double() - refd).abs().max() / refd.abs().max().clamp_min(1e-30)
        assert err <= 1e-5, f"fused backward+push differs: {err}"
        dist.barrier()
        # (3) the autograd wrapper with int32 ids
        t.recv_vals.zero_()
        torch.cuda.synchronize(); dist.barrier()
        tile_a, fm2_a = S.lookup_fm2_autograd(t, ids.int())
        assert torch.equal(tile_a, tile) and torch.equal(fm2_a, fm2)
        ((tile_a * d_tile).sum() + (fm2_a.reshape(-1) * d_fm2).sum()).backward()
        t.finish_push()
        got = t.received_to_dense()
        err = (got.double() - refd).abs().max() / refd.abs().max().clamp_min(1e-30)
        assert err <= 1e-5, f"autograd sharded lookup differs: {err}"
        dist.barrier()
        # (4) fused dense(1) head through autograd: lin = tile @ w, d_w and the pushed rows
        t.recv_vals.zero_()
        torch.cuda.synchronize(); dist.barrier()
        wl = torch.randn((F * D, 1), device=dev, generator=gi).requires_grad_()
        d_lin = torch.randn((B,), device=dev, generator=gi)
        f_l, lin = S.lookup_fm2_linear_autograd(t, ids, wl)
        assert torch.equal(f_l, fm2)
        ((f_l.reshape(-1) * d_fm2).sum() + (lin.reshape(-1) * d_lin).sum()).backward()
        t.finish_push()
        rg_lin = ops.embed_fm2_bwd(tile, (d_lin[:, None] * wl.detach().reshape(1, -1)).reshape(B, F, D).contiguous(), d_fm2)
        ref_l = R.exchange_reference(t.local_rows, t.field_row_offset, ids, rg_lin)
        got = t.received_to_dense()
        err = (got.double() - ref_l).abs().max() / ref_l.abs().max().clamp_min(1e-30)
        assert err <= 1e-5, f"fused-head sharded backward differs: {err}"
        dw_ref = (d_lin.double()[:, None] * tile.double().reshape(B, F * D)).sum(0)
        err = (wl.grad.double().reshape(-1) - dw_ref).abs().max() / dw_ref.abs().max().clamp_min(1e-30)
        assert err <= 1e-5, f"fused-head d_w differs: {err}"
        dist.barrier()
        # ---- owner-side Adam on the received entries (SURVEY 8f.3 on the sharded path): two steps, lazy and TF-dense
        from recalgorithm_b200 import optim
        for lazy in (True, False):
            t.weight.copy_(S.full_to_shard(full, rank, world))
            opt = optim.ShardedTableAdam(t, lr=0.01, lazy=lazy)
            w64, m64, v64 = t.weight.double().clone(), torch.zeros_like(t.weight, dtype=torch.float64), torch.zeros_like(t.weight, dtype=torch.float64)
            for step in (1, 2):
                ids_s = (torch.rand((B, F), device=dev, generator=gi) * (rt[None, :] + 3)).long() - 1
                rg = torch.randn((B, F, D), device=dev, generator=gi)
                t.push_grads(ids_s, rg)
                gd = R.exchange_reference(t.local_rows, t.field_row_offset, ids_s, rg)
                touched = torch.zeros((t.local_rows,), dtype=torch.bool, device=dev)
                for src in range(world):
                    n = int(t.recv_counts[src])
                    touched[t.recv_rows[src, :n]] = True
                opt.step()
                assert opt.last_unique_rows() == int(touched.sum()) and bool((opt._slot == -1).all())
                w64, m64, v64 = R.adam_reference(w64, m64, v64, gd, touched, step, 0.01, lazy)
                for name, a, b_ in (("var", t.weight, w64), ("m", opt.m, m64), ("v", opt.v, v64)):
                    err = (a.double() - b_).abs().max() / b_.abs().max().clamp_min(1e-30)
                    assert err <= 1e-5, f"sharded adam lazy={lazy} step {step} {name}: {err}"
            dist.barrier()
    print(f"SHARDED_OK rank={rank}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
