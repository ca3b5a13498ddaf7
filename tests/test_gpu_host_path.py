"""The host-buffer entry (AggregationStage.forward_host: pinned host boundary tensors, H2D of the head logits on a copy
stream, zero-copy gather of the kept proposals' feature rows, chunk-pipelined, one D2H per chunk) must return exactly
what forward() returns for the same tensors resident on the device -- same kernels, same inputs, bit-identical."""
import pytest
import torch

pytestmark = pytest.mark.gpu

F, LF, C, D = 8, 3, 7, 256
HW = [(24, 24), (12, 12), (6, 6)]


def _synth(B, seed, logits_cl=False):
    fmt = torch.channels_last if logits_cl else torch.contiguous_format
    g = torch.Generator(device="cuda").manual_seed(seed)
    n = B * F
    out = dict(reg=[], obj=[], cls=[], f_cls=[], f_reg=[], f_edge=[])
    for (h, w) in HW:
        xy = torch.rand(n, 2, h, w, generator=g, device="cuda") * 2 - 0.5
        wh = torch.randn(n, 2, h, w, generator=g, device="cuda") * 0.7 + 1.0
        out["reg"].append(torch.cat([xy, wh], 1).half().contiguous(memory_format=fmt))
        out["obj"].append((torch.randn(n, 1, h, w, generator=g, device="cuda") * 2 - 3).half().contiguous(memory_format=fmt))
        out["cls"].append((torch.randn(n, C, h, w, generator=g, device="cuda") * 2 - 3).half().contiguous(memory_format=fmt))
        for k in ("f_cls", "f_reg", "f_edge"):
            out[k].append(torch.randn(n, D, h, w, generator=g, device="cuda").half().contiguous(memory_format=torch.channels_last))
    return out


@pytest.mark.parametrize("graph", [True, False])
@pytest.mark.parametrize("logits_cl", [False, True, "rows"])
@pytest.mark.parametrize("B,chunk", [(5, 2), (4, 4)])
def test_forward_host_equals_forward(B, chunk, logits_cl, graph):
    from tscd_b200 import ops, selection, stage, weights
    cfg = stage.StageConfig(num_classes=C, selection=selection.SelectionConfig(mode="A", pre_k=200, top_k=12, nms_thresh=0.75))
    st = stage.AggregationStage(cfg, weights.random_state_dict(C, D, seed=5))
    dev = _synth(B, 31, logits_cl is True)
    host = {k: [torch.empty(t.shape, dtype=t.dtype, pin_memory=True,
                                memory_format=torch.channels_last if (k.startswith("f_") or logits_cl is True) else torch.contiguous_format).copy_(t) for t in v]
        for k, v in dev.items()}
    if logits_cl == "rows":          # the drop-in head's fused layout: rows read in place over PCIe, objectness plane copied
        packed = ops.pack_head(ops.HeadViews.from_levels(dev["reg"], dev["obj"], dev["cls"], ops.AnchorSpec(HW)))
        torch.cuda.synchronize()
        for k in ("reg", "obj", "cls"):
            del host[k]
        host["rows"], host["objp"] = [packed._keep[0].cpu().pin_memory()], [packed._keep[1].cpu().pin_memory()]
    te = torch.cat([weights.timing_signal_1d(torch.arange(LF), 256)] * B, 0)

    an = ops.AnchorSpec(HW)
    head = ops.HeadViews.from_levels(dev["reg"], dev["obj"], dev["cls"], an)
    feats = tuple(ops.view_levels(dev[k]) for k in ("f_cls", "f_reg", "f_edge"))
    want = []
    # reference run chunk by chunk (the CAFM state is per forward() call, clips are independent)
    for c0 in range(0, B, chunk):
        nc = min(chunk, B - c0)
        f0, f1 = c0 * F, (c0 + nc) * F
        h = ops.HeadViews.from_levels([t[f0:f1] for t in dev["reg"]], [t[f0:f1] for t in dev["obj"]], [t[f0:f1] for t in dev["cls"]], an)
        fv = tuple(ops.view_levels([t[f0:f1] for t in dev[k]]) for k in ("f_cls", "f_reg", "f_edge"))
        out = st.forward(h, fv, torch.float16, te[c0 * LF:(c0 + nc) * LF], nc, F, LF)
        r, o = st.to_lists(out, nc, LF)
        want += list(zip(r, o))
    del head, feats

    for _ in range(2):       # twice: staging buffers are reused across calls
        res, res_ori, h2d, d2h = st.forward_host(host, HW, te.pin_memory(), B, F, LF, chunk_clips=chunk, graph=graph)
        assert len(res) == B * LF and len(res_ori) == B * LF
        n_det = 0
        for (wr, wo), gr, go in zip(want, res, res_ori):
            assert (wr is None) == (gr is None)
            if wr is None:
                continue
            assert gr.device.type == "cpu" and go.device.type == "cpu"
            assert torch.equal(gr, wr.cpu()) and torch.equal(go, wo.cpu())
            n_det += len(gr)
        assert n_det > 0
        feat_bytes = sum(t.numel() * t.element_size() for k in ("f_cls", "f_reg", "f_edge") for t in host[k])
        if logits_cl == "rows":
            head_bytes = sum(t.numel() * t.element_size() for t in host["objp"])
            rows_bytes = sum(t.numel() * t.element_size() for t in host["rows"])
            assert head_bytes < h2d < head_bytes + rows_bytes // 2 + feat_bytes // 4   # 200 of 756 rows + the kept feature rows
        else:
            head_bytes = sum(t.numel() * t.element_size() for k in ("reg", "obj", "cls") for t in host[k])
            assert head_bytes < h2d < head_bytes + feat_bytes // 4      # only the kept rows of the feature planes crossed PCIe
        assert d2h > 0
    if logits_cl is True:
        # optional mode: only the objectness plane is copied, K1 / K3 read the survivors' class / regression rows in place
        res2, ori2, h2d_zc, _ = st.forward_host(host, HW, te.pin_memory(), B, F, LF, chunk_clips=chunk, zero_copy_logits=True, graph=graph)
        assert h2d_zc < h2d
        for gr, go, g2, o2 in zip(res, res_ori, res2, ori2):
            assert (gr is None) == (g2 is None)
            if gr is not None:
                assert torch.equal(gr, g2) and torch.equal(go, o2)


def test_forward_host_rejects_pageable_memory():
    from tscd_b200 import selection, stage, weights
    cfg = stage.StageConfig(num_classes=C, selection=selection.SelectionConfig(mode="A", pre_k=200, top_k=12))
    st = stage.AggregationStage(cfg, weights.random_state_dict(C, D, seed=5))
    dev = _synth(1, 3)
    host = {k: [t.cpu() for t in v] for k, v in dev.items()}
    te = weights.timing_signal_1d(torch.arange(LF), 256)
    with pytest.raises(RuntimeError):
        st.forward_host(host, HW, te, 1, F, LF)


@pytest.mark.parametrize("graph", [True, False])
def test_forward_host_two_calls_in_flight(graph):
    """forward_host_submit / forward_host_collect: two calls in flight on two slots (different input sets, separate buffers and
    graphs) return exactly what the synchronous calls return; a busy slot cannot be resubmitted."""
    from tscd_b200 import ops, selection, stage, weights
    B, chunk = 4, 2
    cfg = stage.StageConfig(num_classes=C, selection=selection.SelectionConfig(mode="A", pre_k=200, top_k=12, nms_thresh=0.75))
    st = stage.AggregationStage(cfg, weights.random_state_dict(C, D, seed=5))
    te = torch.cat([weights.timing_signal_1d(torch.arange(LF), 256)] * B, 0).pin_memory()
    hosts = []
    for seed in (31, 47):
        dev = _synth(B, seed)
        packed = ops.pack_head(ops.HeadViews.from_levels(dev["reg"], dev["obj"], dev["cls"], ops.AnchorSpec(HW)))
        torch.cuda.synchronize()
        h = {k: [torch.empty(t.shape, dtype=t.dtype, pin_memory=True, memory_format=torch.channels_last).copy_(t) for t in dev[k]]
             for k in ("f_cls", "f_reg", "f_edge")}
        h["rows"], h["objp"] = [packed._keep[0].cpu().pin_memory()], [packed._keep[1].cpu().pin_memory()]
        hosts.append(h)
    want = [st.forward_host(h, HW, te, B, F, LF, chunk_clips=chunk, graph=graph) for h in hosts]
    assert sum(len(r) for r in want[0][0] if r is not None) > 0
    for rounds in range(3):
        t0 = st.forward_host_submit(hosts[0], HW, te, B, F, LF, chunk_clips=chunk, graph=graph, slot=0)
        t1 = st.forward_host_submit(hosts[1], HW, te, B, F, LF, chunk_clips=chunk, graph=graph, slot=1)
        with pytest.raises(RuntimeError):
            st.forward_host_submit(hosts[0], HW, te, B, F, LF, chunk_clips=chunk, graph=graph, slot=0)
        for ticket, w in ((t0, want[0]), (t1, want[1])):
            res, ori, _, _ = st.forward_host_collect(ticket)
            for gr, go, wr, wo in zip(res, ori, w[0], w[1]):
                assert (gr is None) == (wr is None)
                if gr is not None:
                    assert torch.equal(gr, wr) and torch.equal(go, wo)
