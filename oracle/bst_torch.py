"""TEST INFRASTRUCTURE: float64 torch restatement of BST/transformer_layer.py:6-79 whose autograd gives the backward oracle
(the reference has no backward code: TF autodiff).  Forward equals oracle.layers_np.bst_transformer_fwd, including the
float32 mask add that makes masked query rows attend uniformly; its gradient is the identity, as the gradient of TF's add is."""
import math

import torch


def _ln(x, beta, gamma):
    mean = x.mean(dim=(1, 2), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(1, 2), keepdim=True)
    inv = torch.rsqrt(var + 1e-12) * gamma
    return x * inv + (beta - mean * inv)


def bst_transformer(q, k, v, keys_length, p, heads, use_position_embedding=True):
    B, T, d = q.shape
    if use_position_embedding:
        q = q + p["position_embedding"][None, :T]
        k = k + p["position_embedding"][None, :T]
    masked = (torch.arange(T)[None, :] >= keys_length[:, None])[:, :, None]
    outs = []
    for h in range(heads):
        Q, K, V = q @ p["w_q"][h], k @ p["w_k"][h], v @ p["w_v"][h]
        S = (Q @ K.transpose(1, 2)) / math.sqrt(d)
        S32 = (S.detach().float() + torch.tensor(float(-2 ** 32 + 1), dtype=torch.float32)).double()
        S = torch.where(masked, S32 + (S - S.detach()), S)          # value S32 exactly, gradient of the add = 1
        outs.append(torch.softmax(S, dim=-1) @ V)
    net = _ln(torch.cat(outs, dim=-1) @ p["w_o"] + q, p["ln1_beta"], p["ln1_gamma"])
    f = net @ p["dense_kernel"] + p["dense_bias"]
    f = 0.5 * (1 + 0.01) * f + 0.5 * (1 - 0.01) * f.abs()
    return _ln(f + net, p["ln2_beta"], p["ln2_gamma"])


def bst_transformer_bwd(q, k, v, keys_length, p, heads, g, use_position_embedding=True):
    """numpy in / numpy out: returns (out, dq, dk, dv, {param: grad})."""
    tq, tk, tv = (torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (q, k, v))
    tp = {n: torch.tensor(a, dtype=torch.float64, requires_grad=True) for n, a in p.items()}
    out = bst_transformer(tq, tk, tv, torch.as_tensor(keys_length), tp, heads, use_position_embedding)
    out.backward(torch.tensor(g, dtype=torch.float64))
    zero = lambda t: t.grad.numpy() if t.grad is not None else torch.zeros_like(t).numpy()
    return out.detach().numpy(), zero(tq), zero(tk), zero(tv), {n: zero(t) for n, t in tp.items()}
