"""Host-side mirror of the reference's layer interface (the drop-in boundary).

The reference has no package API: each ``model_fn`` calls plain Python functions that create their
parameters by side effect with ``tf.get_variable`` inside ``tf.variable_scope`` blocks.  The functions
below keep those names, argument meanings, variable names/shapes and error behaviour, so that a
``model_fn`` body ported to this engine reads like the reference's own:

    with variable_scope("cross_part"):                 # DCN/dcn.py:156-160
        cross_vec = concat_all
        for i in range(params["num_cross_layer"]):
            cross_vec = cross_layer(x0=concat_all, xl=cross_vec, index=i)

Each call is one launch of the matching sm_100a kernel (through recalgorithm_b200.autograd); there
is no other implementation behind these names.
"""
from __future__ import annotations

import contextlib
import math
from typing import Dict, List, Optional, Tuple

import torch

from . import autograd

AUTO_REUSE = "AUTO_REUSE"


class VariableStore:
    """Name-keyed parameter registry mirroring TF1 variable scopes (``tf.get_variable`` semantics:
    create on first use with the default glorot-uniform initializer, reuse afterwards)."""

    def __init__(self, device="cuda", seed: Optional[int] = None):
        self.device = torch.device(device)
        self.vars: Dict[str, torch.nn.Parameter] = {}
        self.scope: List[str] = []
        self.gen = torch.Generator(device="cpu")
        if seed is not None:
            self.gen.manual_seed(seed)

    def full_name(self, name: str) -> str:
        return "/".join(self.scope + [name])

    def get_variable(self, name: str, shape, initializer=None) -> torch.nn.Parameter:
        full = self.full_name(name)
        shape = tuple(int(s) for s in shape)
        if full in self.vars:
            v = self.vars[full]
            if tuple(v.shape) != shape:
                raise ValueError(f"Trying to share variable {full}, but specified shape {shape} and found shape {tuple(v.shape)}.")
            return v
        if initializer is None:                         # tf.get_variable default: glorot_uniform_initializer
            if len(shape) >= 2:
                rf = 1
                for s in shape[:-2]:
                    rf *= s
                fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
            else:
                fan_in = fan_out = shape[0] if shape else 1
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            data = (torch.rand(shape, generator=self.gen) * 2 - 1) * lim
        elif callable(initializer):
            data = initializer(shape)
        else:
            data = torch.as_tensor(initializer, dtype=torch.float32).reshape(shape)
        v = torch.nn.Parameter(data.to(torch.float32).to(self.device).contiguous())
        self.vars[full] = v
        return v

    def assign(self, values: Dict[str, "torch.Tensor"]):
        """Inject weights by their TF variable names (parity is defined on injected weights)."""
        for k, val in values.items():
            t = torch.as_tensor(val, dtype=torch.float32).to(self.device).contiguous()
            if k in self.vars:
                if tuple(self.vars[k].shape) != tuple(t.shape):
                    raise ValueError(f"{k}: shape {tuple(t.shape)} != {tuple(self.vars[k].shape)}")
                self.vars[k].data.copy_(t)
            else:
                self.vars[k] = torch.nn.Parameter(t)

    def parameters(self):
        return list(self.vars.values())


_default_store: Optional[VariableStore] = None


def default_store() -> VariableStore:
    global _default_store
    if _default_store is None:
        _default_store = VariableStore()
    return _default_store


def set_default_store(store: VariableStore) -> VariableStore:
    global _default_store
    _default_store = store
    return store


@contextlib.contextmanager
def variable_scope(name: str, reuse=None):
    st = default_store()
    st.scope.append(name)
    try:
        yield
    finally:
        st.scope.pop()


def get_variable(name, shape, dtype=None, initializer=None):
    return default_store().get_variable(name, shape, initializer)


# --------------------------------------------------------------------------------------------------- DCN
def cross_layer(x0: torch.Tensor, xl: torch.Tensor, index: int) -> torch.Tensor:
    """dcn cross layer -- same signature as DCN/cross_layer.py:4.  Variables ``wl_{index}``, ``bl_{index}`` of shape
    (d, 1), default (glorot-uniform) initialised -- the bias is NOT zero-initialised in the reference (:18-19)."""
    dimension = int(x0.shape[-1])
    wl = get_variable(name=f"wl_{index}", shape=(dimension, 1))
    bl = get_variable(name=f"bl_{index}", shape=(dimension, 1))
    return autograd.cross_stack(x0, wl.reshape(1, dimension), bl.reshape(1, dimension), xl=None if xl is x0 else xl)


def cross_network(x0: torch.Tensor, num_cross_layer: int) -> torch.Tensor:
    """The whole loop of DCN/dcn.py:157-160 (``for i: cross_vec = cross_layer(x0, cross_vec, i)``) in ONE launch;
    creates exactly the variables the loop would create."""
    dimension = int(x0.shape[-1])
    if num_cross_layer == 0:
        return x0
    ws = [get_variable(name=f"wl_{i}", shape=(dimension, 1)) for i in range(num_cross_layer)]
    bs = [get_variable(name=f"bl_{i}", shape=(dimension, 1)) for i in range(num_cross_layer)]
    w = torch.cat([t.reshape(1, dimension) for t in ws], 0)
    b = torch.cat([t.reshape(1, dimension) for t in bs], 0)
    return autograd.cross_stack(x0, w, b)


# --------------------------------------------------------------------------------------------------- xDeepFM
def cin_layer(x0: torch.Tensor, xk: torch.Tensor, hk_1, index: int, return_pooled: bool = False):
    """xdeepfm CIN layer -- same signature as xDeepFM/cin_layer.py:4.  x0 (B,m,D), xk (B,hk,D) -> (B,hk_1,D).
    ``hk_1`` may arrive as a *string* (the reference splits a comma flag, xdeepfm.py:253).  Variable
    ``cin_layer_{index}_filter`` of shape (1, hk*m, hk_1).  ``return_pooled`` additionally returns sum over D
    (the ``tf.reduce_sum(x, axis=-1)`` of xdeepfm.py:173) from the same kernel."""
    hk_1 = int(hk_1)
    m = int(x0.shape[1])
    hk = int(xk.shape[1])
    filters = get_variable(name=f"cin_layer_{index}_filter", shape=(1, hk * m, hk_1))
    return autograd.cin(x0, xk, filters[0], want_pooled=return_pooled)


# --------------------------------------------------------------------------------------------------- DIN
def din_attention(query: torch.Tensor, keys: torch.Tensor, keys_length: torch.Tensor, is_softmax: bool = False):
    """DIN attention unit -- same signature as DIN/din_attention.py:4.  Dense layers ``f1_att`` (4H->64, relu),
    ``f2_att`` (64->32, relu), ``f3_att`` (32->1) with AUTO_REUSE: variables <name>/kernel, <name>/bias (zeros)."""
    H = int(query.shape[-1])
    params = []
    for name, (fi, fo) in (("f1_att", (4 * H, 64)), ("f2_att", (64, 32)), ("f3_att", (32, 1))):
        with variable_scope(name, reuse=AUTO_REUSE):
            params.append(get_variable("kernel", (fi, fo)))
            params.append(get_variable("bias", (fo,), initializer=lambda s: torch.zeros(s)))
    return autograd.din_attention(query, keys, keys_length.to(torch.int64), *params, is_softmax=is_softmax)


# --------------------------------------------------------------------------------------------------- FiBiNET
def senet(input: torch.Tensor, embedding_dim: int, reduction_ratio: int) -> torch.Tensor:
    """SENET -- same signature as FiBiNET/senet.py:4.  NB the reference reduces from ``embedding_dim``
    (``reduction_dim = embedding_dim // reduction_ratio``, :18), not from the field count."""
    F = int(input.shape[1])
    reduction_dim = embedding_dim // reduction_ratio
    assert reduction_dim < embedding_dim, "reduction_dim must be less than embedding_dim"
    w1 = get_variable(name="senet_w1", shape=(F, reduction_dim))
    w2 = get_variable(name="senet_w2", shape=(reduction_dim, F))
    return autograd.senet(input, w1, w2)


def bilinear_interaction_layer(input: torch.Tensor, embedding_dim: int, type: str, name: str) -> torch.Tensor:
    """Bilinear interaction -- same signature as FiBiNET/bilinear_interaction_layer.py:5.  Output is
    (B, (F-1)(F-2)/2, K): the reference enumerates ``combinations(range(F-1), 2)`` (:24,29,33)."""
    F = int(input.shape[1])
    if type == "all":
        w = get_variable(name=f"{name}_w_all", shape=(embedding_dim, embedding_dim))
    elif type == "each":
        w = get_variable(name=f"{name}_w_each", shape=(F - 1, embedding_dim, embedding_dim))
    elif type == "interaction":
        w = get_variable(name=f"{name}_w_interaction", shape=(F * (F - 1) // 2, embedding_dim, embedding_dim))
    else:
        raise ValueError(f"Bilinear Interaction type must be in ['all','each','interaction'], got '{type}'")
    return autograd.bilinear(input, w, type)


# --------------------------------------------------------------------------------------------------- DeepFM
def fm_second_order(tables: autograd.EmbeddingTables, ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """The lookup + FM second-order block of DeepFM/deepfm.py:184-200 in one kernel: per-field ids (B,F) ->
    (fields_embeddings as a (B,F,K) tile, fm_second_order_logit (B,1))."""
    return autograd.lookup_fm2(tables, ids)


# --------------------------------------------------------------------------------------------------- NFM / FwFM / AFM (8f.4)
def bi_interaction(tables: autograd.EmbeddingTables, ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """The lookup + bi-interaction pooling block of NFM/nfm.py:155-168 in one kernel: per-field ids (B,F) ->
    (fields_embeddings (B,F,K), nfm (B,K)) -- the vector the reference then feeds to batch-norm / dropout / the DNN."""
    return autograd.lookup_bi(tables, ids)


def fwfm_second_order(fields_embeddings: torch.Tensor) -> torch.Tensor:
    """FwFM/fwfm.py:145-158: creates ``fields_pair_strength/fields_pair_strength_weight`` (F(F-1)/2,) and returns
    ``fwfm_second_order_logit`` (B,1) for the (B,F,K) field embeddings."""
    F = int(fields_embeddings.shape[1])
    with variable_scope("fields_pair_strength"):
        r = get_variable("fields_pair_strength_weight", shape=(F * (F - 1) // 2,))
    return autograd.fwfm(fields_embeddings, r)


def afm_attention(fields_embeddings: torch.Tensor, embedding_dim: int, attention_factor: int) -> torch.Tensor:
    """AFM/afm.py:152-186: attention-weighted sum of the pairwise hadamard products, (B,K).  Creates
    ``attention_part/attention_{w,b,h}`` with the reference's shapes; ``p`` and the final matmul (afm.py:187-188) stay
    with the caller."""
    with variable_scope("attention_part"):
        w = get_variable(name="attention_w", shape=(embedding_dim, attention_factor))
        b = get_variable(name="attention_b", shape=(attention_factor,))
        h = get_variable(name="attention_h", shape=(attention_factor, 1))
    return autograd.afm(fields_embeddings, w, b, h)


# --------------------------------------------------------------------------------------------------- BST (8f.4)
def bst_transformer(queries: torch.Tensor, keys: torch.Tensor, values: torch.Tensor, keys_length: torch.Tensor, heads: int,
                    index: int, max_length: int, use_position_embedding: bool = True) -> torch.Tensor:
    """Transformer block -- same signature as BST/transformer_layer.py:6.  Variables carry the names TF gives them inside
    the caller's scope (``transformer_part`` in BST/bst.py:184): ``position_embedding`` (shared by every block),
    ``w_{q,k,v,o}_{index}``, ``LayerNorm[/_n]/{beta,gamma}`` and ``dense[/_n]/{kernel,bias}`` with TF's uniquifying suffixes
    for block ``index`` (two layer norms and one dense per block)."""
    d = int(queries.shape[-1])
    suffix = lambda base, n: base if n == 0 else f"{base}_{n}"
    ones, zeros = (lambda s: torch.ones(s)), (lambda s: torch.zeros(s))
    p = {"position_embedding": get_variable(name="position_embedding", shape=(max_length, d)),
         "w_q": get_variable(name=f"w_q_{index}", shape=(heads, d, d)), "w_k": get_variable(name=f"w_k_{index}", shape=(heads, d, d)),
         "w_v": get_variable(name=f"w_v_{index}", shape=(heads, d, d)), "w_o": get_variable(name=f"w_o_{index}", shape=(heads * d, d))}
    with variable_scope(suffix("LayerNorm", 2 * index)):
        p["ln1_beta"] = get_variable("beta", (d,), initializer=zeros); p["ln1_gamma"] = get_variable("gamma", (d,), initializer=ones)
    with variable_scope(suffix("dense", index)):
        p["dense_kernel"] = get_variable("kernel", (d, d)); p["dense_bias"] = get_variable("bias", (d,), initializer=zeros)
    with variable_scope(suffix("LayerNorm", 2 * index + 1)):
        p["ln2_beta"] = get_variable("beta", (d,), initializer=zeros); p["ln2_gamma"] = get_variable("gamma", (d,), initializer=ones)
    return autograd.bst_transformer(queries, keys, values, keys_length.to(torch.int64), p, heads, max_length, use_position_embedding)


# --------------------------------------------------------------------------------------------------- DIN loss-side term
def mini_batch_aware_regularization(embedding_tensors, l2_lambda: float) -> torch.Tensor:
    """DIN/din.py:254-257 (SURVEY parity note 7): ``l2_lambda * tf.nn.l2_loss(concat(tensors, -1)) / batch`` -- an L2 on the
    looked-up ACTIVATIONS, so its gradient ``l2_lambda * e / B`` reaches the tables through the lookup backward
    (``d_tile`` of ctr_embed_fm2_bwd / ctr_bag_lookup_bwd) like any other upstream gradient.  Torch plumbing, no kernel."""
    x = torch.cat(list(embedding_tensors), dim=-1)
    return l2_lambda * 0.5 * x.pow(2).sum() / x.shape[0]


# --------------------------------------------------------------------------------------------------- FFM
def ffm_table_dim(num_fields: int, embedding_dim: int) -> int:
    """Row width of the id-major FFM table: (F-1)*K padded up to the next width the fused 128-bit gather takes (a power of
    two in 4..128) -- the reference's default config (7 fields, K = 8: 48 floats) becomes 64, the padding is zeros."""
    d = (num_fields - 1) * embedding_dim
    p = 4
    while p < d:
        p *= 2
    if p > 128:
        raise ValueError(f"(F-1)*embedding_dim = {d} exceeds the 128-float row of the fused lookup")
    return p


def ffm_second_order(tables: autograd.EmbeddingTables, ids: torch.Tensor, embedding_dim: int) -> torch.Tensor:
    """FFM/ffm.py:128-160 in two kernels: `tables` holds, per field, rows of (F-1)*K floats (zero padded to
    ``ffm_table_dim(F, K)``) -- the reference's ``{name}_embedding`` variable of shape (F-1, |V_i|, K) stored id-major
    (``EmbeddingTables([...], dim=ffm_table_dim(F, K))``; `ffm_table_from_reference` converts) -- looked up by the fused
    gather, then the field-aware pair sum.  Returns (B,1).  Single-valued fields only (the reference's mean-combined
    ``manual_tag_list`` bag is outside this entry point: look it up with ``ctr_bag_lookup_*`` per slot)."""
    B, F = ids.shape
    d = (F - 1) * embedding_dim
    if tables.dim != ffm_table_dim(F, embedding_dim):
        raise ValueError(f"tables.dim must be ffm_table_dim(F, K) = {ffm_table_dim(F, embedding_dim)}, got {tables.dim}")
    tile = autograd.lookup(tables, ids)                       # (B, F, padded (F-1)*K)
    if tables.dim != d:
        tile = tile[:, :, :d].contiguous()
    return autograd.ffm(tile.reshape(B, F, F - 1, embedding_dim))


def ffm_table_from_reference(embedding_variables) -> torch.Tensor:
    """[(F-1, |V_i|, K) per field] (the reference's variables, ffm.py:129-136) -> the (sum |V_i|, ffm_table_dim) id-major table."""
    F = embedding_variables[0].shape[0] + 1
    K = embedding_variables[0].shape[2]
    t = torch.cat([e.permute(1, 0, 2).reshape(e.shape[1], -1) for e in embedding_variables], dim=0)
    pad = ffm_table_dim(F, K) - t.shape[1]
    return (torch.nn.functional.pad(t, (0, pad)) if pad else t).contiguous()
