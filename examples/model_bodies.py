"""The bodies of the reference's model_fns between "inputs looked up" and "logit", rewritten on this repo's host API.

Each function follows its reference file line by line (cited); the interaction layers keep the reference's names, scopes and
signatures (recalgorithm_b200.layers), the lookup is the fused kernel (autograd.lookup / lookup_fm2), and the dense tail --
out of scope for the library -- is plain torch with variables created under the same TF names through
`dense()` below (glorot kernel, zero bias, like tf.layers.dense).  batch-norm / dropout / Dice of the tails are left out.

tests/test_gpu_model_bodies.py runs every body forward + backward once and checks that gradients reach the tables
(IndexedSlices) and every variable.
"""
from __future__ import annotations

import torch

from recalgorithm_b200 import autograd
from recalgorithm_b200 import layers as L


def dense(x: torch.Tensor, units: int, activation=None, use_bias: bool = True, name: str = "dense") -> torch.Tensor:
    """tf.layers.dense: variables <scope>/<name>/kernel (in, units), <scope>/<name>/bias (units,)."""
    with L.variable_scope(name):
        kernel = L.get_variable("kernel", (x.shape[-1], units))
        y = x @ kernel
        if use_bias:
            y = y + L.get_variable("bias", (units,), initializer=lambda s: torch.zeros(s))
    return activation(y) if activation is not None else y


def deepfm_logit(tables: autograd.EmbeddingTables, ids: torch.Tensor, fm_first_order_logit: torch.Tensor, hidden_units=(64, 32)):
    """DeepFM/deepfm.py:178-214.  fm_first_order_logit comes from feature_column.indicator_dense (deepfm.py:180-181)."""
    fields_embeddings, fm_second_order_logit = L.fm_second_order(tables, ids)            # :184-200, one kernel
    with L.variable_scope("fm_deep"):                                                      # :203-211
        net = fields_embeddings.reshape(ids.shape[0], -1)                                 # tf.concat(fields_embeddings, axis=1)
        for i, unit in enumerate(hidden_units):
            net = dense(net, unit, activation=torch.relu, name=f"dense_{i}" if i else "dense")
        deep_logit = dense(net, 1, name="deep_logit")
    return fm_first_order_logit + fm_second_order_logit + deep_logit                      # tf.add_n(...), :214


def dcn_logit(dense_input: torch.Tensor, category_input: torch.Tensor, num_cross_layer: int = 3, hidden_units=(64, 32)):
    """DCN/dcn.py:147-169.  category_input: (B, sum d) flat output of the lookup."""
    concat_all = torch.cat([dense_input, category_input], dim=-1)                         # :155
    with L.variable_scope("cross_part"):                                                  # :157-160
        cross_vec = concat_all
        for i in range(num_cross_layer):
            cross_vec = L.cross_layer(x0=concat_all, xl=cross_vec, index=i)
    with L.variable_scope("dnn_part"):                                                    # :162-165
        dnn_vec = concat_all
        for i, unit in enumerate(hidden_units):
            dnn_vec = dense(dnn_vec, unit, activation=torch.relu, name=f"dnn_dense_{i}")
    with L.variable_scope("output_part"):                                                 # :167-169
        return dense(torch.cat([cross_vec, dnn_vec], dim=-1), 1)


def xdeepfm_logit(dense_input: torch.Tensor, x0: torch.Tensor, cin_layer_feature_maps=("16", "16"), hidden_units=(64, 32)):
    """xDeepFM/xdeepfm.py:152-185.  x0: (B, m, D) tile; layer widths arrive as strings like in the reference (:253)."""
    B = x0.shape[0]
    category_input = x0.reshape(B, -1)
    with L.variable_scope("linear_part"):                                                 # :160-163
        linear_vec = torch.cat([dense_input, category_input], dim=-1)
        linear_logit = dense(linear_vec, 1)
    with L.variable_scope("cin_part"):                                                    # :166-175
        xk, p_plus = x0, []
        for i, features_map_num in enumerate(cin_layer_feature_maps):
            xk, pooled = L.cin_layer(x0, xk, features_map_num, i + 1, return_pooled=True)   # pooled = reduce_sum(x, axis=-1), fused
            p_plus.append(pooled)
        cin_logit = dense(torch.cat(p_plus, dim=-1), 1, use_bias=False)
    with L.variable_scope("dnn_part"):                                                    # :178-182
        dnn_vec = linear_vec
        for i, unit in enumerate(hidden_units):
            dnn_vec = dense(dnn_vec, unit, activation=torch.relu, name=f"dense_{i}")
        dnn_logit = dense(dnn_vec, 1, use_bias=False, name="dnn_logit")
    return linear_logit + cin_logit + dnn_logit                                           # :185


def din_logit(dense_input, category_input, target_input, sequnence_input, sequnence_length, use_softmax=False, hidden_units=(64, 32)):
    """DIN/din.py:199-238 (Dice / PReLU / batch-norm of the fcn tail left out)."""
    with L.variable_scope("attention_part"):                                              # :216-218
        attention_output = L.din_attention(target_input, sequnence_input, sequnence_length, is_softmax=use_softmax)
    concat_all = torch.cat([dense_input, category_input, target_input, attention_output], dim=-1)   # :221
    with L.variable_scope("fcn"):                                                         # :224-238
        net = concat_all
        for i, unit in enumerate(hidden_units):
            net = torch.relu(dense(net, unit, name=f"dense_{i}" if i else "dense"))
        return dense(net, 1, name="logit"), attention_output


def fibinet_logit(dense_input, category_input, embedding_dim: int, reduction_ratio: int = 2, bilinear_interaction_type: str = "all",
                  hidden_units=(64, 32)):
    """FiBiNET/fibinet.py:156-199.  category_input: (B, F, K)."""
    with L.variable_scope("linear_part"):                                                 # :166-168
        linear_logit = dense(dense_input, 1)
    with L.variable_scope("senet_part"):                                                  # :171-174
        senet_output = L.senet(category_input, embedding_dim=embedding_dim, reduction_ratio=reduction_ratio)
    with L.variable_scope("bilinear_interaction_part"):                                   # :177-187
        bi_orginal = L.bilinear_interaction_layer(category_input, embedding_dim=embedding_dim, type=bilinear_interaction_type, name="orginal")
        bi_senet = L.bilinear_interaction_layer(senet_output, embedding_dim=embedding_dim, type=bilinear_interaction_type, name="senet")
        bi_total = torch.cat([bi_orginal, bi_senet], dim=-1).reshape(category_input.shape[0], -1)
    with L.variable_scope("dnn_part"):                                                    # :189-197
        net = bi_total
        for i, unit in enumerate(hidden_units):
            net = dense(net, unit, activation=torch.relu, name=f"dense_{i}" if i else "dense")
        fibinet = dense(net, 1, name="logit")
    return linear_logit + fibinet                                                         # :199
