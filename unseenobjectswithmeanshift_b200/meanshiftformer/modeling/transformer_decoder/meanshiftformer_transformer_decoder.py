"""Mean-shift transformer decoders - mirror of the reference's
modeling/transformer_decoder/meanshiftformer_transformer_decoder.py (layers :27-315, MLP :329-341,
MeanShiftTransformerDecoder :343-695, PretrainedMeanShiftTransformerDecoder :697-1048).

Same class names, constructor keywords, ``from_config`` keys, registry and ``state_dict`` layout
(``transformer_cross_attention_layers.{i}.meanshift_attn.*`` ... ``mask_embed.layers.{j}.*``), so
reference checkpoints load with ``strict=True``. What differs is how ``forward`` runs:

* batch-first [B, len, C] buffers throughout; heads are addressed by strides (no transposes);
* the sine position embedding is a cached [S, C] table added to the keys' input, not a
  [B, C, H, W] tensor rebuilt per call;
* keys/values of all layers that share a feature level are projected in one GEMM before the
  sequential loop (they do not depend on the queries);
* cross/self attention run in the streaming vMF kernel (csrc/vmf_attention.cu): no [B*h, Q, S]
  score, weight or additive-mask tensors;
* the attention mask is 1 bit per (query, key) shared by all heads (csrc/mask_head.cu), with the
  reference's "row blocks everything -> attend everywhere" rule (:618) carried as a per-row flag.
"""
import logging
import os
from typing import Optional

import torch
from torch import Tensor, nn
from torch.nn import functional as F

from .... import ops
from ....d2compat import Conv2d, c2_xavier_fill, configurable
from .attention_util import MeanShiftAttention
from .maskformer_transformer_decoder import TRANSFORMER_DECODER_REGISTRY
from .position_encoding import PositionEmbeddingSine
from .transformer import _get_activation_fn

# keys+values of one feature level are pre-projected for all its layers when they fit this budget
_KV_PRECOMPUTE_BYTES = 4 << 30


def _xavier(module):
    for p in module.parameters():
        if p.dim() > 1:
            nn.init.xavier_uniform_(p)


class _AttentionLayerBase(nn.Module):
    def with_pos_embed(self, tensor, pos: Optional[Tensor]):
        return tensor if pos is None else tensor + pos


class SelfAttentionLayer(_AttentionLayerBase):
    """Vanilla softmax self-attention layer (reference :27-87); selected only when
    USE_MEANSHIFT_SELF_ATTENTION is False, which no UOIS config does. Plain torch."""

    def __init__(self, d_model, nhead, dropout=0.0, activation="relu", normalize_before=False):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.norm = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)
        self.activation = _get_activation_fn(activation)
        self.normalize_before = normalize_before
        _xavier(self)

    def forward(self, tgt, tgt_mask=None, tgt_key_padding_mask=None, query_pos=None):
        if self.normalize_before:
            t2 = self.norm(tgt)
            q = k = self.with_pos_embed(t2, query_pos)
            return tgt + self.dropout(self.self_attn(q, k, value=t2, attn_mask=tgt_mask,
                                                     key_padding_mask=tgt_key_padding_mask)[0])
        q = k = self.with_pos_embed(tgt, query_pos)
        t2 = self.self_attn(q, k, value=tgt, attn_mask=tgt_mask, key_padding_mask=tgt_key_padding_mask)[0]
        return self.norm(tgt + self.dropout(t2))


class CrossAttentionLayer(_AttentionLayerBase):
    """Vanilla softmax cross-attention layer (reference :90-145); unused by UOIS configs. Plain torch."""

    def __init__(self, d_model, nhead, dropout=0.0, activation="relu", normalize_before=False):
        super().__init__()
        self.multihead_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.norm = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)
        self.activation = _get_activation_fn(activation)
        self.normalize_before = normalize_before
        _xavier(self)

    def forward(self, tgt, memory, memory_mask=None, memory_key_padding_mask=None, pos=None, query_pos=None):
        if self.normalize_before:
            t2 = self.norm(tgt)
            t2 = self.multihead_attn(query=self.with_pos_embed(t2, query_pos), key=self.with_pos_embed(memory, pos),
                                     value=memory, attn_mask=memory_mask,
                                     key_padding_mask=memory_key_padding_mask)[0]
            return tgt + self.dropout(t2)
        t2 = self.multihead_attn(query=self.with_pos_embed(tgt, query_pos), key=self.with_pos_embed(memory, pos),
                                 value=memory, attn_mask=memory_mask, key_padding_mask=memory_key_padding_mask)[0]
        return self.norm(tgt + self.dropout(t2))


class MeanShiftSelfAttentionLayer(_AttentionLayerBase):
    """Reference :148-203. Stand-alone ``forward`` keeps the reference's seq-first signature."""

    def __init__(self, d_model, nhead, dropout=0.0, activation="relu", normalize_before=False):
        super().__init__()
        self.self_attn = MeanShiftAttention(d_model, nhead, dropout=dropout)
        self.norm = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)
        self.activation = _get_activation_fn(activation)
        self.normalize_before = normalize_before
        _xavier(self)

    def forward_post(self, tgt, tgt_mask=None, tgt_key_padding_mask=None, query_pos=None):
        q = k = self.with_pos_embed(tgt, query_pos)
        t2 = self.self_attn(q, k, value=tgt, attn_mask=tgt_mask, key_padding_mask=tgt_key_padding_mask,
                            need_weights=False)[0]
        return self.norm(tgt + self.dropout(t2))

    def forward_pre(self, tgt, tgt_mask=None, tgt_key_padding_mask=None, query_pos=None):
        t2 = self.norm(tgt)
        q = k = self.with_pos_embed(t2, query_pos)
        t2 = self.self_attn(q, k, value=t2, attn_mask=tgt_mask, key_padding_mask=tgt_key_padding_mask,
                            need_weights=False)[0]
        return tgt + self.dropout(t2)

    def forward(self, tgt, tgt_mask=None, tgt_key_padding_mask=None, query_pos=None):
        if self.normalize_before:
            return self.forward_pre(tgt, tgt_mask, tgt_key_padding_mask, query_pos)
        return self.forward_post(tgt, tgt_mask, tgt_key_padding_mask, query_pos)


class MeanShiftCrossAttentionLayer(_AttentionLayerBase):
    """Reference :206-272."""

    def __init__(self, d_model, nhead=1, dropout=0.0, activation="relu", layer_normalize_before=False):
        super().__init__()
        self.meanshift_attn = MeanShiftAttention(d_model, nhead, dropout=dropout)
        self.norm = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)
        self.activation = _get_activation_fn(activation)
        self.normalize_before = layer_normalize_before
        _xavier(self)

    def forward_pre(self, tgt, memory, memory_mask=None, memory_key_padding_mask=None, pos=None, query_pos=None):
        t2 = self.norm(tgt)
        t2 = self.meanshift_attn(query=self.with_pos_embed(t2, query_pos), key=self.with_pos_embed(memory, pos),
                                 value=memory, attn_mask=memory_mask, key_padding_mask=memory_key_padding_mask,
                                 need_weights=False)[0]
        return tgt + self.dropout(t2)

    def forward_post(self, tgt, memory, memory_mask=None, memory_key_padding_mask=None, pos=None, query_pos=None):
        t2 = self.meanshift_attn(query=self.with_pos_embed(tgt, query_pos), key=self.with_pos_embed(memory, pos),
                                 value=memory, attn_mask=memory_mask, key_padding_mask=memory_key_padding_mask,
                                 need_weights=False)[0]
        return self.norm(tgt + self.dropout(t2))

    def forward(self, tgt, memory, memory_mask=None, memory_key_padding_mask=None, pos=None, query_pos=None):
        if self.normalize_before:
            return self.forward_pre(tgt, memory, memory_mask, memory_key_padding_mask, pos, query_pos)
        return self.forward_post(tgt, memory, memory_mask, memory_key_padding_mask, pos, query_pos)


class FFNLayer(nn.Module):
    """Reference :275-315."""

    def __init__(self, d_model, dim_feedforward=2048, dropout=0.0, activation="relu", normalize_before=False):
        super().__init__()
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm = nn.LayerNorm(d_model)
        self.activation = _get_activation_fn(activation)
        self.normalize_before = normalize_before
        _xavier(self)

    def forward(self, tgt):
        def ffn(t):   # ops.dense: tensor-core GEMM in inference AND under autograd (fp32 training, DenseFunction)
            relu = self.activation is F.relu
            h = ops.dense(t, self.linear1.weight, self.linear1.bias, relu=relu)
            if not relu:
                h = self.activation(h)
            return ops.dense(self.dropout(h), self.linear2.weight, self.linear2.bias)

        if self.normalize_before:
            return tgt + self.dropout(ffn(self.norm(tgt)))
        return self.norm(tgt + self.dropout(ffn(tgt)))


class MLP(nn.Module):
    """Reference :329-341."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))

    def forward(self, x):
        for i, layer in enumerate(self.layers):
            x = ops.dense(x, layer.weight, layer.bias, relu=i < self.num_layers - 1)
        return x


class _MeanShiftDecoderBase(nn.Module):
    """Shared body of the two registered decoders; they differ only in ``_NUM_LEVELS`` (3 vs 1,
    reference :494 vs :848)."""

    _version = 2
    _NUM_LEVELS = 3

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        version = local_metadata.get("version", None)
        if version is None or version < 2:  # reference :348-369: static_query -> query_feat
            renamed = False
            for k in list(state_dict.keys()):
                if k.startswith(prefix) and "static_query" in k:
                    state_dict[k.replace("static_query", "query_feat")] = state_dict.pop(k)
                    renamed = True
            if renamed:
                logging.getLogger(__name__).warning(
                    f"Weight format of {self.__class__.__name__} have changed! "
                    "Please upgrade your models. Applying automatic conversion now ...")
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)

    @configurable
    def __init__(self, in_channels, mask_classification=True, *, num_classes: int, hidden_dim: int, num_queries: int,
                 nheads: int, dim_feedforward: int, dec_layers: int, pre_norm: bool, mask_dim: int,
                 enforce_input_project: bool, use_meanshift_cross_attention: bool, disable_attention_mask: bool,
                 use_meanshift_self_attention: bool, decoder_block_norm: bool):
        super().__init__()
        assert mask_classification, "Only support mask classification model"
        self.mask_classification = mask_classification
        self.pe_layer = PositionEmbeddingSine(hidden_dim // 2, normalize=True)
        self.num_heads = nheads
        self.num_layers = dec_layers
        self.pre_norm = pre_norm
        self.use_meanshift_seeds = False  # hard-coded in the reference (:424, :778)
        # Reference behaviour (True): every prediction, also in eval mode, carries its full-resolution mask logits in
        # ``aux_outputs``. The eval branch of the META_ARCH never reads them (pretrained_meanshiftformer_model.py:
        # 335-378), so the wrappers of this package set it False: intermediate layers then compute their masks only on
        # the next layer's key grid (``_lean_masks``) and ``aux_outputs[i]["pred_masks"]`` holds those low-resolution
        # logits. Training (grad enabled) always produces the full masks.
        self.eval_aux_masks = True
        self.use_meanshift_cross_attention = use_meanshift_cross_attention
        self.disable_attention_mask = disable_attention_mask
        self.use_meanshift_self_attention = use_meanshift_self_attention
        self.decoder_block_norm = decoder_block_norm
        self.transformer_self_attention_layers = nn.ModuleList()
        self.transformer_cross_attention_layers = nn.ModuleList()
        self.transformer_ffn_layers = nn.ModuleList()
        for _ in range(self.num_layers):
            sa = MeanShiftSelfAttentionLayer if use_meanshift_self_attention else SelfAttentionLayer
            self.transformer_self_attention_layers.append(
                sa(d_model=hidden_dim, nhead=nheads, dropout=0.0, normalize_before=pre_norm))
            if use_meanshift_cross_attention:
                self.transformer_cross_attention_layers.append(MeanShiftCrossAttentionLayer(
                    d_model=hidden_dim, nhead=nheads, dropout=0.0, layer_normalize_before=pre_norm))
            else:
                self.transformer_cross_attention_layers.append(CrossAttentionLayer(
                    d_model=hidden_dim, nhead=nheads, dropout=0.0, normalize_before=pre_norm))
            self.transformer_ffn_layers.append(FFNLayer(d_model=hidden_dim, dim_feedforward=dim_feedforward,
                                                        dropout=0.0, normalize_before=pre_norm))
        self.decoder_norm = nn.LayerNorm(hidden_dim)
        self.num_queries = num_queries
        self.query_feat = nn.Embedding(num_queries, hidden_dim)
        self.query_embed = nn.Embedding(num_queries, hidden_dim)
        self.num_feature_levels = self._NUM_LEVELS
        self.level_embed = nn.Embedding(self.num_feature_levels, hidden_dim)
        self.input_proj = nn.ModuleList()
        for _ in range(self.num_feature_levels):
            if in_channels != hidden_dim or enforce_input_project:
                self.input_proj.append(Conv2d(in_channels, hidden_dim, kernel_size=1))
                c2_xavier_fill(self.input_proj[-1])
            else:
                self.input_proj.append(nn.Sequential())
        if self.mask_classification:
            self.class_embed = nn.Linear(hidden_dim, num_classes + 1)
        self.mask_embed = MLP(hidden_dim, hidden_dim, mask_dim, 3)

    @classmethod
    def from_config(cls, cfg, in_channels, mask_classification):
        mf = cfg.MODEL.MASK_FORMER
        assert mf.DEC_LAYERS >= 1
        return {
            "in_channels": in_channels,
            "mask_classification": mask_classification,
            "num_classes": cfg.MODEL.SEM_SEG_HEAD.NUM_CLASSES,
            "hidden_dim": mf.HIDDEN_DIM,
            "num_queries": mf.NUM_OBJECT_QUERIES,
            "nheads": mf.NHEADS,
            "dim_feedforward": mf.DIM_FEEDFORWARD,
            "dec_layers": mf.DEC_LAYERS - 1,  # reference :529: one "layer" is the learnable-query prediction
            "pre_norm": mf.PRE_NORM,
            "enforce_input_project": mf.ENFORCE_INPUT_PROJ,
            "mask_dim": cfg.MODEL.SEM_SEG_HEAD.MASK_DIM,
            "use_meanshift_cross_attention": mf.USE_MEANSHIFT_CROSS_ATTENTION,
            "disable_attention_mask": mf.DISABLE_MEANSHIFT_ATTENTION_MASK,
            "use_meanshift_self_attention": mf.USE_MEANSHIFT_SELF_ATTENTION,
            "decoder_block_norm": mf.DECODER_BLOCK_NORM,
        }

    # ------------------------------------------------------------------ prediction heads
    def _heads(self, out, mask_features, target_size, need_mask, dec=None, lean_features=None):
        """out [B,Q,C] -> (class logits [B,Q,K+1], mask logits [B,Q,h,w], bits, row_open).
        ``dec`` = decoder_norm(out) when the caller already has it (fused into the previous GEMM's epilogue).
        ``lean_features``: the mask features already resampled to ``target_size`` (see ``eval_aux_masks``) - the mask
        logits are then computed at that resolution only."""
        if dec is None:
            dec = self.decoder_norm(out)
        logits = ops.dense(dec, self.class_embed.weight, self.class_embed.bias)   # N = K + 1: padded to 32 columns
        embed = self.mask_embed(dec)
        if lean_features is not None:
            return (logits,) + self._lean_masks(embed, lean_features, target_size)
        if torch.is_grad_enabled() and (embed.requires_grad or mask_features.requires_grad):
            masks = ops.mask_logits_autograd(embed, mask_features)  # training (row f4)
        else:
            masks = ops.mask_logits(embed, mask_features)
        bits = row_open = None
        if need_mask:  # the attention mask is a constant of the graph (reference :680 detaches it)
            bits, row_open = ops.mask_to_attn_bits(masks.detach(), target_size)
        return logits, masks, bits, row_open

    @staticmethod
    def _lean_masks(embed, lean_features, target_size):
        """Inference without auxiliary full-resolution masks: interpolate(einsum(e, F)) == einsum(e, interpolate(F))
        (both linear), so an intermediate layer's attention-mask bits need the mask logits only on the NEXT layer's key
        grid: 300 / 1200 / 4800 pixels instead of 19200, and 157 MB of mask features are not re-read per layer.
        -> (low-resolution logits [B,Q,ht,wt], bits, row_open)."""
        masks = ops.mask_logits(embed, lean_features)
        bits, row_open = ops.mask_to_attn_bits(masks, target_size)
        return masks, bits, row_open

    def forward_prediction_heads(self, output, mask_features, attn_mask_target_size):
        """Reference :660-682 / :1012-1035 (seq-first ``output`` [Q,B,C]); returns the reference's
        (outputs_class, outputs_mask, bool attn_mask [B*heads, Q, S] or None)."""
        logits, masks, bits, _ = self._heads(output.transpose(0, 1).contiguous(), mask_features,
                                             attn_mask_target_size, not self.disable_attention_mask)
        attn_mask = None
        if bits is not None:
            S = int(attn_mask_target_size[0]) * int(attn_mask_target_size[1])
            attn_mask = ops.unpack_attn_bits(bits, torch.ones_like(bits[..., 0]), S, self.num_heads)
        return logits, masks, attn_mask

    # ------------------------------------------------------------------ forward
    def forward(self, x, mask_features, mask=None, _teacher=None):
        """Reference :540-658 / :894-1010. ``_teacher`` is a test hook (tests/test_gpu_config2.py): a callable
        ``(layer, out, bits, row_open) -> (out, bits, row_open)`` run at the top of every layer, so that a parity test
        can feed each layer the ORACLE's inputs (teacher forcing) instead of letting a flipped mask bit of an earlier
        layer decide what later layers see."""
        assert len(x) == self.num_feature_levels
        del mask  # reference :548 / :900
        if not (self.use_meanshift_cross_attention and self.use_meanshift_self_attention) or self.pre_norm:
            raise NotImplementedError(
                "the CUDA path implements the configuration every UOIS YAML selects: post-norm with "
                "mean-shift cross- and self-attention")
        B = x[0].shape[0]
        C = self.query_feat.weight.shape[1]
        H, L = self.num_heads, self.num_feature_levels
        hd = C // H
        dev = x[0].device
        mask_features = mask_features.float().contiguous()
        # training: the attention core and the mask head run through their autograd Functions
        # (native forward + backward), every other layer through torch (cuBLAS / ATen, autograd-capable)
        train = torch.is_grad_enabled()

        # ---- per-level memory (keys' input = src + pos, values' input = src), batch-first [B,S,C]
        sizes = [tuple(x[l].shape[-2:]) for l in range(L)]
        key_in = [None] * L   # src + positional table, built only where the add cannot be folded (below)

        class _Src:
            """src[l], computed on first use: the packed K / V path folds input_proj into the projections and then
            never needs the [B, S, C] map (315 MB per image at 480x640)"""
            def __init__(self):
                self.cache = {}

            def __getitem__(self_, l):
                if l not in self_.cache:
                    xf = x[l].float()
                    proj = self.input_proj[l]
                    if isinstance(proj, nn.Conv2d):
                        bias = proj.bias + self.level_embed.weight[l]
                        if not torch.is_grad_enabled() and ops.conv1x1_supported(xf, proj.weight):
                            s_ = ops.conv1x1(xf, proj.weight, bias, tokens_out=True)      # NCHW in, [B,S,C] out
                        else:  # e.g. the pixel decoder's token-major maps seen through a transposed view
                            s_ = ops.dense(xf.flatten(2).transpose(1, 2), proj.weight.flatten(1), bias)
                    else:
                        s_ = xf.flatten(2).transpose(1, 2) + self.level_embed.weight[l]
                    self_.cache[l] = s_
                return self_.cache[l]

        src = _Src()

        # ---- keys / values: independent of the queries, so project them up front per level
        layers_of = [[i for i in range(self.num_layers) if i % L == l] for l in range(L)]
        kv = {}

        # default (MSM_PACKED_KV=0 switches it off): K / V projections write the attention kernel's operand images
        # instead of fp32 rows (DESIGN.md sections 4.2, 4.3)
        packed_kv = ops.packed_kv_enabled() and not train and hd == 32

        def project_kv(level, layer_ids):
            attn = [self.transformer_cross_attention_layers[i].meanshift_attn for i in layer_ids]
            tag = "kv%d_%s" % (level, "_".join(map(str, layer_ids)))
            if packed_kv:
                S_l = sizes[level][0] * sizes[level][1]
                cat = lambda name, ts: ops.cached_cat(self, name, ts)  # noqa: E731
                wk = cat(tag + "wk", [a.in_proj_weight[C:2 * C] for a in attn])
                bk = cat(tag + "bk", [a.in_proj_bias[C:2 * C] for a in attn])
                wv = cat(tag + "wv", [a.in_proj_weight[2 * C:] for a in attn])
                bv = cat(tag + "bv", [a.in_proj_bias[2 * C:] for a in attn])
                # a fresh buffer per forward (a graph's own under capture: several graphs may be in flight at once)
                images, per_layer = ops.packed_kv_alloc(len(layer_ids), B, H, S_l, dev)
                proj = self.input_proj[level]
                if (isinstance(proj, nn.Conv2d) and proj.weight.shape[1] % 32 == 0 and proj.weight.shape[1] < C
                        and os.environ.get("MSM_FOLD_V", "1") == "1"):
                    #  input_proj folded into both projections: they read the 64-channel map instead of a
                    # 256-channel one that is then never built - 4x fewer FLOPs and input bytes.
                    #   values = (W_in x + b_in + level_embed) W_v^T + b_v = x (W_v W_in)^T + const
                    #   keys   = (W_in x + b_in + level_embed + pos) W_k^T + b_k
                    #          = x (W_k W_in)^T + const + ty[y] + tx[x]     (the sine embedding is separable:
                    #            channels [0, C/2) depend on y only, [C/2, C) on x only - position_encoding.py)
                    w_in = proj.weight.flatten(1)
                    const = proj.bias + self.level_embed.weight[level]
                    x_tok = x[level].float().contiguous()     # the channel-major map itself: no transposed copy
                    if S_l % 4:
                        x_tok = x_tok.flatten(2).transpose(1, 2).contiguous()
                    fold_w = lambda w_: (w_.double() @ w_in.double()).float().contiguous()  # noqa: E731
                    fold_b = lambda w_, b_: (w_.double() @ const.double() + b_.double()).float().contiguous()  # noqa: E731
                    wv_f = ops.cached_value(self, tag + "wvf", [wv, proj.weight], lambda: fold_w(wv))
                    bv_f = ops.cached_value(self, tag + "bvf", [wv, bv, proj.bias, self.level_embed.weight],
                                            lambda: fold_b(wv, bv))
                    ops.linear_packed_kv(x_tok, wv_f, bv_f, images, B, S_l, C, 1)
                    if os.environ.get("MSM_FOLD_K", "1") == "1":
                        wk_f = ops.cached_value(self, tag + "wkf", [wk, proj.weight], lambda: fold_w(wk))
                        bk_f = ops.cached_value(self, tag + "bkf", [wk, bk, proj.bias, self.level_embed.weight],
                                                lambda: fold_b(wk, bk))
                        ty, tx = self.pe_layer.tables(sizes[level][0], sizes[level][1], dev)
                        npf = ty.shape[1]
                        pos = ops.cached_value(
                            self, tag + "postab_%dx%d" % sizes[level], [wk],
                            lambda: ((ty.double() @ wk[:, :npf].double().t()).float().contiguous(),
                                     (tx.double() @ wk[:, npf:].double().t()).float().contiguous()))
                        ops.linear_packed_kv(x_tok, wk_f, bk_f, images, B, S_l, C, 0, pos=pos)
                    else:
                        if key_in[level] is None:
                            key_in[level] = src[level] + self.pe_layer.table(sizes[level][0], sizes[level][1], dev)
                        ops.linear_packed_kv(key_in[level].contiguous(), wk, bk, images, B, S_l, C, 0)
                else:
                    if key_in[level] is None:
                        key_in[level] = src[level] + self.pe_layer.table(sizes[level][0], sizes[level][1], dev)
                    ops.linear_packed_kv(key_in[level].contiguous(), wk, bk, images, B, S_l, C, 0)
                    ops.linear_packed_kv(src[level].contiguous(), wv, bv, images, B, S_l, C, 1)
                for j, i in enumerate(layer_ids):
                    kv[i] = (ops.PackedKV(images[j * per_layer:(j + 1) * per_layer], B, H, S_l), None)
                return
            # the cache holds detached copies: under autograd the concatenation has to stay part of the graph
            cat = (lambda name, ts: torch.cat(ts, 0)) if train else (lambda name, ts: ops.cached_cat(self, name, ts))
            wk = cat(tag + "wk", [a.in_proj_weight[C:2 * C] for a in attn])
            bk = cat(tag + "bk", [a.in_proj_bias[C:2 * C] for a in attn])
            wv = cat(tag + "wv", [a.in_proj_weight[2 * C:] for a in attn])
            bv = cat(tag + "bv", [a.in_proj_bias[2 * C:] for a in attn])
            table = self.pe_layer.table(sizes[level][0], sizes[level][1], dev)
            if (os.environ.get("MSM_FOLD_POS", "0") == "1" and not torch.is_grad_enabled() and ops.tc_linear_enabled()
                    and ops.linear_supported(src[level], wk) and table.numel() * len(layer_ids) * 4 <= (64 << 20)):
                # (src + pos) Wk^T = src Wk^T + (pos Wk^T): the positional half is a cached [S, n*C] row bias.
                # Opt-in: at M = 38400 rows the strided row-bias reads slow the HBM-bound epilogue by more than the
                # add kernel costs (4.90 -> 5.14 ms per step with both folds on, B200)
                tab = ops.cached_value(self, tag + "pos", [table, wk], lambda: F.linear(table, wk).contiguous())
                K = ops.linear_fused(src[level], wk, bk, rowbias=tab)  # [B,S,n*C]
            else:
                if key_in[level] is None:
                    key_in[level] = src[level] + table
                K = ops.dense(key_in[level], wk, bk)  # [B,S,n*C]
            V = ops.dense(src[level], wv, bv)
            for j, i in enumerate(layer_ids):
                kv[i] = (K[..., j * C:(j + 1) * C], V[..., j * C:(j + 1) * C])

        for l in range(L):
            S = sizes[l][0] * sizes[l][1]
            if layers_of[l] and 8 * B * S * C * len(layers_of[l]) <= _KV_PRECOMPUTE_BYTES:
                project_kv(l, layers_of[l])

        def heads_view(t):  # [B,len,C] (row stride may exceed C) -> [B,H,len,hd] view
            return t.unflatten(-1, (H, hd)).permute(0, 2, 1, 3)

        def attention(q_, k_, v_, bits_=None, row_open_=None):  # [B,len,C] projections -> [B,Q,C]
            if isinstance(k_, ops.PackedKV):
                o_ = torch.empty(B, self.num_queries, C, device=dev, dtype=torch.float32)
                ops.vmf_attention_packed(heads_view(q_), k_, blocked_bits=bits_, row_open=row_open_, out=heads_view(o_))
                return o_
            if train:
                o4 = ops.vmf_attention_autograd(heads_view(q_), heads_view(k_), heads_view(v_), blocked_bits=bits_,
                                                row_open=row_open_)
                return o4.permute(0, 2, 1, 3).reshape(B, self.num_queries, C)
            o_ = torch.empty(B, self.num_queries, C, device=dev, dtype=torch.float32)
            ops.vmf_attention(heads_view(q_), heads_view(k_), heads_view(v_), blocked_bits=bits_, row_open=row_open_,
                              out=heads_view(o_))
            return o_

        query_pos = self.query_embed.weight.unsqueeze(0)  # [1,Q,C]
        out = self.query_feat.weight.unsqueeze(0).expand(B, -1, -1).contiguous()
        need_mask = not self.disable_attention_mask

        # opt-in (MSM_L2_PERSIST=1; measured slower on the B200): the mask features are re-read by every
        # prediction head call below - keep as much of them as the device allows resident in L2
        l2_window = ops.l2_persist_enabled() and not train
        if l2_window:
            ops.l2_persist(mask_features)

        lean = (not train and not self.eval_aux_masks and need_mask and _teacher is None
                and any(tuple(sz) != tuple(mask_features.shape[-2:]) for sz in sizes))
        lean_cache = {}

        def lean_features(size):
            """mask features resampled (bilinear, align_corners=False, as :675) to a key grid, once per forward"""
            if not lean or tuple(size) == tuple(mask_features.shape[-2:]):
                return None
            key = tuple(size)
            if key not in lean_cache:
                lean_cache[key] = ops.resample_bilinear(mask_features, key)
            return lean_cache[key]

        predictions_class, predictions_mask = [], []
        logits, masks, bits, row_open = self._heads(out, mask_features, sizes[0], need_mask,
                                                    lean_features=lean_features(sizes[0]))
        predictions_class.append(logits)
        predictions_mask.append(masks)

        # Inference fast path: every post-norm residual block ends in ONE GEMM launch whose epilogue does
        # + residual, LayerNorm (and for the FFN block F.normalize + the heads' decoder_norm), and
        # in_proj(tgt + query_pos) becomes in_proj(tgt) + a cached [Q, N] row-bias table.
        fused = (not torch.is_grad_enabled() and C % 32 == 0 and C <= 256 and self.mask_classification
                 and isinstance(self.decoder_norm, nn.LayerNorm) and ops.tc_linear_enabled()
                 and all(isinstance(l.norm, nn.LayerNorm) for l in self.transformer_ffn_layers)
                 and self.transformer_ffn_layers[0].linear1.weight.shape[0] % 32 == 0)
        qpos = self.query_embed.weight
        # MSM_DECODER_FUSION: 0 = separate add / LayerNorm kernels; 1 (default) = the query_pos row bias is folded
        # into the in-projections; 2 = residual + LayerNorm (+ normalise + decoder_norm) GEMM epilogues as well -
        # measured slower at B*Q = 800 rows (one 256-column CTA per 128 rows = 7 CTAs; 5.5 vs 6.3 ms per step)
        level = int(os.environ.get("MSM_DECODER_FUSION", "1"))
        fused = fused and level > 0

        # One cluster kernel per layer for everything between the cross-attention and the mask einsum
        # (csrc/decoder_block.cu): the configuration of every UOIS YAML (hidden 256, 8 heads, FFN 2048, relu, post-norm)
        block = (fused and ops.decoder_block_enabled() and C == 256 and H == 8 and self.num_queries <= 128
                 and all(f.linear1.weight.shape[0] == 2048 and f.activation is F.relu for f in self.transformer_ffn_layers)
                 and self.class_embed.weight.shape[0] <= 32 and self.mask_embed.num_layers == 3
                 and self.mask_embed.layers[2].weight.shape[0] == 256)
        q_carry = None   # the next layer's cross-attention query projection, produced by the previous layer's kernel

        def tq_table(j):
            aj = self.transformer_cross_attention_layers[j].meanshift_attn
            return ops.cached_value(self, f"tq{j}", [qpos, aj.in_proj_weight],
                                    lambda: F.linear(qpos, aj.in_proj_weight[:C]).contiguous())

        for i in range(self.num_layers):
            lvl = i % L
            if _teacher is not None:
                out, bits, row_open = _teacher(i, out, bits, row_open)
                q_carry = None   # the carried projection belongs to the state the teacher just replaced
            if i not in kv:
                project_kv(lvl, [i])
            K, V = kv.pop(i)
            ca = self.transformer_cross_attention_layers[i]
            sl = self.transformer_self_attention_layers[i]
            ffn = self.transformer_ffn_layers[i]
            a = ca.meanshift_attn
            sa = sl.self_attn
            if block:
                q = q_carry if q_carry is not None else ops.linear_fused(out, a.in_proj_weight[:C], a.in_proj_bias[:C],
                                                                         rowbias=tq_table(i))
                if isinstance(K, ops.PackedKV):
                    o = attention(q, K, V, bits, row_open)
                else:
                    o = torch.empty(B, self.num_queries, C, device=dev, dtype=torch.float32)
                    ops.vmf_attention(heads_view(q), heads_view(K), heads_view(V), blocked_bits=bits, row_open=row_open,
                                      out=heads_view(o))
                del K, V
                nxt = self.transformer_cross_attention_layers[i + 1].meanshift_attn if i + 1 < self.num_layers else None
                mlp = self.mask_embed.layers
                deps = [a.out_proj.weight, sa.in_proj_weight, sa.out_proj.weight, ffn.linear1.weight, ffn.linear2.weight,
                        mlp[0].weight, self.class_embed.weight, mlp[1].weight, mlp[2].weight] + (
                            [nxt.in_proj_weight] if nxt is not None else [])
                blob = ops.cached_value(self, f"dbk{i}", deps, lambda: ops.decoder_block_pack(
                    a.out_proj.weight, sa.in_proj_weight, sa.out_proj.weight, ffn.linear1.weight, ffn.linear2.weight,
                    nxt.in_proj_weight[:C] if nxt is not None else None, mlp[0].weight, self.class_embed.weight,
                    mlp[1].weight, mlp[2].weight))
                tqk = ops.cached_value(self, f"tqk{i}", [qpos, sa.in_proj_weight],
                                       lambda: torch.cat([F.linear(qpos, sa.in_proj_weight[:2 * C]),
                                                          qpos.new_zeros(qpos.shape[0], C)], 1).contiguous())
                bc32 = ops.cached_value(self, "bc32", [self.class_embed.bias], lambda: torch.cat(
                    [self.class_embed.bias.detach(), self.class_embed.bias.new_zeros(32 - self.class_embed.bias.numel())]))
                out, logits32, embed, q_carry = ops.decoder_block(
                    o, out, blob, b_o1=a.out_proj.bias, norm1=ca.norm, b_qkv=sa.in_proj_bias, t_qk=tqk,
                    b_o2=sa.out_proj.bias, norm2=sl.norm, b_f1=ffn.linear1.bias, b_f2=ffn.linear2.bias, norm3=ffn.norm,
                    block_norm=self.decoder_block_norm, normd=self.decoder_norm,
                    b_qn=nxt.in_proj_bias[:C] if nxt is not None else None,
                    t_qn=tq_table(i + 1) if nxt is not None else None, b_m1=mlp[0].bias, b_c32=bc32, b_m2=mlp[1].bias,
                    b_m3=mlp[2].bias)
                logits = logits32[..., :self.class_embed.weight.shape[0]]
                lf = lean_features(sizes[(i + 1) % L]) if i + 1 < self.num_layers else None
                if lf is not None:
                    masks, bits, row_open = self._lean_masks(embed, lf, sizes[(i + 1) % L])
                else:
                    masks = ops.mask_logits(embed, mask_features)
                    bits = row_open = None
                    if need_mask:
                        bits, row_open = ops.mask_to_attn_bits(masks, sizes[(i + 1) % L])
                predictions_class.append(logits)
                predictions_mask.append(masks)
                continue
            if fused and ffn.activation is F.relu:
                # cross-attention (reference :245-260), post-norm
                tq = tq_table(i)
                q = ops.linear_fused(out, a.in_proj_weight[:C], a.in_proj_bias[:C], rowbias=tq)
                if isinstance(K, ops.PackedKV):
                    o = attention(q, K, V, bits, row_open)
                else:
                    o = torch.empty(B, self.num_queries, C, device=dev, dtype=torch.float32)
                    ops.vmf_attention(heads_view(q), heads_view(K), heads_view(V), blocked_bits=bits, row_open=row_open,
                                      out=heads_view(o))
                if level >= 2:
                    out = ops.linear_fused(o, a.out_proj.weight, a.out_proj.bias, residual=out, norm=ca.norm)
                else:
                    out = ops.add_layernorm(out, ops.dense(o, a.out_proj.weight, a.out_proj.bias), ca.norm)
                del K, V
                # self-attention (reference :171-181): q = k = out + query_pos, v = out -> one GEMM, N = 3C
                tqk = ops.cached_value(self, f"tqk{i}", [qpos, sa.in_proj_weight],
                                       lambda: torch.cat([F.linear(qpos, sa.in_proj_weight[:2 * C]),
                                                          qpos.new_zeros(qpos.shape[0], C)], 1).contiguous())
                qkv = ops.linear_fused(out, sa.in_proj_weight, sa.in_proj_bias, rowbias=tqk)
                o = torch.empty(B, self.num_queries, C, device=dev, dtype=torch.float32)
                ops.vmf_attention(heads_view(qkv[..., :C]), heads_view(qkv[..., C:2 * C]), heads_view(qkv[..., 2 * C:]),
                                  out=heads_view(o))
                if level >= 2:
                    out = ops.linear_fused(o, sa.out_proj.weight, sa.out_proj.bias, residual=out, norm=sl.norm)
                    # FFN (reference :300-304), block norm (:637-638) and the heads' decoder_norm (:663)
                    hdn = ops.dense(out, ffn.linear1.weight, ffn.linear1.bias, relu=True)
                    out, dec = ops.linear_fused(hdn, ffn.linear2.weight, ffn.linear2.bias, residual=out, norm=ffn.norm,
                                                l2_normalize=self.decoder_block_norm, norm2=self.decoder_norm)
                else:
                    out = ops.add_layernorm(out, ops.dense(o, sa.out_proj.weight, sa.out_proj.bias), sl.norm)
                    # FFN (reference :300-304), block norm (:637-638) and the heads' decoder_norm (:663)
                    t2 = ops.dense(ops.dense(out, ffn.linear1.weight, ffn.linear1.bias, relu=True),
                                   ffn.linear2.weight, ffn.linear2.bias)
                    out, dec = ops.add_layernorm(out, t2, ffn.norm, l2_normalize=self.decoder_block_norm,
                                                 norm2=self.decoder_norm)
                logits, masks, bits, row_open = self._heads(
                    out, mask_features, sizes[(i + 1) % L], need_mask, dec=dec,
                    lean_features=lean_features(sizes[(i + 1) % L]) if i + 1 < self.num_layers else None)
                predictions_class.append(logits)
                predictions_mask.append(masks)
                continue
            # cross-attention (reference :245-260), post-norm
            q = ops.dense(out + query_pos, a.in_proj_weight[:C], a.in_proj_bias[:C])
            o = attention(q, K, V, bits, row_open)
            out = ca.norm(out + ops.dense(o, a.out_proj.weight, a.out_proj.bias))
            del K, V
            # self-attention (reference :171-181): q = k = out + query_pos, v = out
            a = sa
            qk = ops.dense(out + query_pos, a.in_proj_weight[:2 * C], a.in_proj_bias[:2 * C])
            v = ops.dense(out, a.in_proj_weight[2 * C:], a.in_proj_bias[2 * C:])
            o = attention(qk[..., :C], qk[..., C:], v)
            out = sl.norm(out + ops.dense(o, a.out_proj.weight, a.out_proj.bias))
            # FFN (reference :300-304) and block norm (:637-638)
            out = ffn(out)
            if self.decoder_block_norm:
                out = F.normalize(out, dim=-1)
            logits, masks, bits, row_open = self._heads(
                out, mask_features, sizes[(i + 1) % L], need_mask,
                lean_features=lean_features(sizes[(i + 1) % L]) if i + 1 < self.num_layers else None)
            predictions_class.append(logits)
            predictions_mask.append(masks)

        assert len(predictions_class) == self.num_layers + 1
        if l2_window:
            ops.l2_persist(None)
        return {
            "pred_logits": predictions_class[-1],
            "pred_masks": predictions_mask[-1],
            "aux_outputs": self._set_aux_loss(predictions_class if self.mask_classification else None,
                                              predictions_mask),
        }

    @torch.jit.unused
    def _set_aux_loss(self, outputs_class, outputs_seg_masks):
        if self.mask_classification:
            return [{"pred_logits": a, "pred_masks": b} for a, b in zip(outputs_class[:-1], outputs_seg_masks[:-1])]
        return [{"pred_masks": b} for b in outputs_seg_masks[:-1]]


@TRANSFORMER_DECODER_REGISTRY.register()
class MeanShiftTransformerDecoder(_MeanShiftDecoderBase):
    """Reference :343-695: three feature levels, cycled ``i % 3`` (ResNet-50 configs)."""
    _NUM_LEVELS = 3


@TRANSFORMER_DECODER_REGISTRY.register()
class PretrainedMeanShiftTransformerDecoder(_MeanShiftDecoderBase):
    """Reference :697-1048: one full-resolution level (UCN RGB-D configs)."""
    _NUM_LEVELS = 1
