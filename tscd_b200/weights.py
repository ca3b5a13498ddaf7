"""Random-init aggregation-stage parameters with the reference's state_dict key names and shapes
(TSCDHead.__init__, yolox/models/tscd_head.py:92-133).  Used by bench.py / smoke() where no
checkpoint is available; real deployments pass `model.head.state_dict()`."""
import math

import torch


def random_state_dict(num_classes: int, dim: int = 256, seed: int = 2024, gen1: bool = False):
    """gen1=True: the gen-1 (YOLOV) MSA head instead -- trans.msa.qkv_*, trans.linear1/2, linear_pred (yolovp_msa.py)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(name, out_f, in_f, bias=True):
        bound = 1.0 / math.sqrt(in_f)                      # nn.Linear default init
        sd[name + ".weight"] = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound
        if bias:
            sd[name + ".bias"] = (torch.rand(out_f, generator=g) * 2 - 1) * bound

    def ln(name, n):
        sd[name + ".weight"] = torch.ones(n)
        sd[name + ".bias"] = torch.zeros(n)

    D = dim
    if gen1:
        lin("trans.msa.qkv_cls", 3 * D, D, False); lin("trans.msa.qkv_reg", 3 * D, D, False)
        lin("trans.linear1", 2 * D, 2 * D); lin("trans.linear2", 4 * D, 4 * D)
        lin("linear_pred", num_classes + 1, 4 * D)
        return sd
    for m in ("agg.", "agg_iou."):
        lin(m + "mca.q_cls_local", D, D, False); lin(m + "mca.kv_cls", 2 * D, D, False)
        lin(m + "mca.q_reg_local", D, D, False); lin(m + "mca.kv_reg", 2 * D, D, False)
        lin(m + "mca.linear", 2 * D, 2 * D); lin(m + "mca.linear_reg", 2 * D, 2 * D)
        lin(m + "linear", 4 * D, 3 * D); lin(m + "linear_obj", 4 * D, 3 * D)
    p = "local_reg_matcher."
    for layer in ("transformer_self_attention_layers.0.self_attn.", "transformer_aware_cross_attention_layers.0.multihead_attn."):
        for n in ("q_reg", "k_reg", "v_reg"):
            lin(p + layer + n, D, D, False)
        sd[p + layer + "position_embedding.weight"] = torch.randn(8, 64, 1, 1, generator=g) * 0.1
        sd[p + layer + "position_embedding.bias"] = torch.zeros(8)
    for layer in ("transformer_self_attention_layers.0.", "transformer_aware_cross_attention_layers.0."):
        ln(p + layer + "norm", D)
        lin(p + layer + "CA.fc.0", 32, 2, False); lin(p + layer + "CA.fc.2", 2, 32, False)
    lin(p + "transformer_ffn_layers.0.linear1", D, D); lin(p + "transformer_ffn_layers.0.linear2", D, D)
    ln(p + "transformer_ffn_layers.0.norm", D)
    lin(p + "absolute_position_embedding", D, 256); lin(p + "edge_feature_embedding", D, D // 4)
    ln(p + "decoder_norm", D)
    lin("fc_reg_matcher", 4 * D, D)
    t = "task_aligned.transformer_cross_attention_layers.0."
    for n in ("q_reg", "k_reg", "v_reg"):
        lin(t + "multihead_attn." + n, 4 * D, 4 * D, False)
    ln(t + "norm", 4 * D); ln("task_aligned.decoder_norm", 4 * D)
    lin("cls_pred", num_classes, 4 * D); lin("matcher_obj_pred", 1, 4 * D); lin("matcher_reg_pred", 4, 4 * D)
    return sd


def timing_signal_1d(index_sequence: torch.Tensor, channels: int = 256, min_timescale=1.0, max_timescale=1.0e4):
    """Sinusoidal time embedding of frame indices (yolox/data/datasets/vid.py:1015-1023)."""
    n = channels // 2
    log_inc = torch.tensor(math.log(max_timescale / min_timescale) / (n - 1))
    inv = min_timescale * torch.exp(torch.arange(0, n) * -log_inc)
    scaled = index_sequence.float().unsqueeze(1) * inv.unsqueeze(0)
    return torch.cat([torch.sin(scaled), torch.cos(scaled)], dim=1)
