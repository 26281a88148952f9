"""Training-mode forward + backward of the YOLOv3 graph on the sm_100a kernels.

Mirrors what the reference runs in ``train.py:401-411`` — ``pred = model(imgs)`` in train mode (BatchNorm with batch
statistics, eps 1e-3 / momentum 0.03; ``Detect`` returning the raw ``[bs,na,ny,nx,no]`` maps, models/yolo.py:110), then
``loss.backward()`` through every Conv block — without autograd graphs or cuDNN:

  forward  per Conv block:  conv (tcgen05 implicit GEMM, identity epilogue) -> bn_stats -> bn_finalize -> bn_act_fwd
  backward per Conv block:  bn_act_bwd (dgamma, dbeta, dy) -> wgrad (bf16 MMA, split over pixels) -> dgrad, which is the
                            SAME conv kernel run on dy with the transposed, tap-flipped weight pack (stride-2 layers:
                            on the zero-stuffed dy), accumulating into the input's gradient through the residual port.

``TrainEngine.forward/backward`` are wrapped in one ``torch.autograd.Function`` so that the reference's
``loss.backward(); optimizer.step()`` work unchanged on the fp32 master parameters (``Model.parameters()``).
Supported layer types in train mode: Conv, Bottleneck, nn.Upsample, Concat, Detect (= yolov3.yaml).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib, ops
from . import train_ops as T
from .tensors import PaddedNHWC, _stream


class _Block:
    """One Conv+BN+SiLU block (models/common.py:57-81) with everything its forward and backward need."""

    __slots__ = ("prefix", "c1", "c2", "k", "s", "x", "y", "a", "res", "upsample", "wf", "wd", "zero_b", "zero_bi", "st",
                 "dw", "dw_tm", "first", "dy", "dy_up", "post_fwd", "pre_bwd")


class TrainEngine:
    def __init__(self, model, n, h, w):
        dev = model.device
        self.model, self.n, self.h, self.w = model, n, h, w
        det = model.detect
        nodes = model.nodes
        for nd in nodes[:-1]:
            if nd.type not in ("Conv", "Bottleneck", "Upsample", "Concat", "SPP"):
                raise NotImplementedError(f"training-mode {nd.type} is not built (yolov3.yaml / yolov3-spp.yaml need "
                                          "Conv/Bottleneck/Upsample/Concat/SPP)")
        P = model.device_params()
        self.P = P
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.sync_bn = bool(getattr(model, "sync_bn", False)) and self.world > 1
        if self.sync_bn:
            self.use_graphs = False  # the per-layer collectives stay eager launches
        self.blocks: list[_Block] = []
        self.keep = []
        self.grad_bufs: dict[int, PaddedNHWC] = {}
        self.scratch: dict[tuple, PaddedNHWC] = {}

        def buf(c, hh, ww, ld=None):
            b = PaddedNHWC.zeros(n, hh, ww, c, device=dev, ld=ld)
            self.keep.append(b)
            return b

        def f32(c):
            t = torch.zeros(c, dtype=torch.float32, device=dev)
            self.keep.append(t)
            return t

        def new_block(prefix, c1, c2, k, s, x, a, res=None, upsample=False, first=False):
            b = _Block()
            b.prefix, b.c1, b.c2, b.k, b.s, b.x, b.a, b.res, b.upsample, b.first = prefix, c1, c2, k, s, x, a, res, upsample, first
            ho, wo = x.h // s, x.w // s
            b.y = buf(c2, ho, wo)
            kk = 1 if first else k
            ci = 32 if first else c1
            b.wf = torch.zeros(ops.cout_pad(c2), kk * kk * ci, dtype=torch.bfloat16, device=dev)
            b.wd = None if first else torch.zeros(ops.cout_pad(c1), k * k * c2, dtype=torch.bfloat16, device=dev)
            b.zero_b = f32(ops.cout_pad(c2))
            b.zero_bi = None if first else f32(ops.cout_pad(c1))
            b.st = {name: f32(c2) for name in ("scale", "shift", "mean", "rstd")}
            sums, dsums, gsums = f32(2 * c2).view(2, c2), f32(2 * c2).view(2, c2), f32(2 * c2).view(2, c2)
            b.st.update(sums=sums, sum=sums[0], sumsq=sums[1], dsums=dsums, dbeta=dsums[0], dgamma=dsums[1], gsums=gsums)
            # wgrad accumulator: [k*k, co, ci] when the tensor-core kernel is in use (vector reductions), else PyTorch's layout
            b.dw_tm = T.wgrad_tap_major(ci)
            b.dw = torch.zeros((kk * kk, c2, ci) if b.dw_tm else (c2, ci, kk, kk), dtype=torch.float32, device=dev)
            b.dy = self._scratch(c2, ho, wo, dev)
            b.dy_up = self._scratch(c2, x.h, x.w, dev, tag="up") if s == 2 else None
            b.post_fwd, b.pre_bwd = [], []  # extra launches after this block's forward / before its backward (SPP pools)
            self.blocks.append(b)
            return b

        # ---- shapes and concat destinations (same zero-copy concat / fused upsample layout as the inference engine)
        shp = {}
        for nd in nodes[:-1]:
            src = [(model.ch, h, w) if s < 0 else shp[s] for s in nd.srcs]
            c0, h0, w0 = src[0]
            if nd.type == "Conv":
                s_ = nd.args[3] if len(nd.args) > 3 else 1
                shp[nd.i] = (nd.c_out, h0 // s_, w0 // s_)
            elif nd.type in ("Bottleneck", "SPP"):
                shp[nd.i] = (nd.c_out, h0, w0)
            elif nd.type == "Upsample":
                shp[nd.i] = (c0, h0 * 2, w0 * 2)
            elif nd.type == "Concat":
                shp[nd.i] = (sum(s[0] for s in src), h0, w0)
        consumers = {}
        for nd in nodes:
            for s in nd.srcs:
                consumers.setdefault(s, []).append(nd.i)
        cat_buf, alias, up_alias = {}, {}, {}
        for nd in nodes[:-1]:
            if nd.type != "Concat":
                continue
            c, hh, ww = shp[nd.i]
            cat = buf(c, hh, ww)
            cat_buf[nd.i] = cat
            off = 0
            for s in nd.srcs:
                cs = shp[s][0]
                sl = cat.slice(off, cs)
                off += cs
                if nodes[s].type == "Upsample":
                    v = nodes[s].srcs[0]
                    assert consumers.get(v) == [s] and consumers.get(s) == [nd.i] and nodes[v].type == "Conv"
                    up_alias[v] = sl
                else:
                    alias[s] = sl

        def out_of(i):
            if i in alias:
                return alias[i]
            c, hh, ww = shp[i]
            return buf(c, hh, ww)

        # ---- lower the graph into Conv blocks
        self.im2col = buf(32, h, w)
        tens = {}
        for nd in nodes[:-1]:
            srcs = [tens[s] if s >= 0 else None for s in nd.srcs]
            base = f"model.{nd.i}"
            reps = [base] if nd.n == 1 else [f"{base}.{j}" for j in range(nd.n)]
            if nd.type == "Conv":
                c1, c2, *rest = nd.args
                k = rest[0] if len(rest) > 0 else 1
                s_ = rest[1] if len(rest) > 1 else 1
                x = srcs[0]
                for ri, r in enumerate(reps):
                    last = ri == len(reps) - 1
                    if x is None:
                        assert c1 == 3 and k == 3 and s_ == 1
                        a = out_of(nd.i) if last else buf(c2, h, w)
                        new_block(r, c1, c2, k, 1, self.im2col, a, first=True)
                    elif last and nd.i in up_alias:
                        a = up_alias[nd.i]
                        new_block(r, c1, c2, k, s_, x, a, upsample=True)
                    else:
                        a = out_of(nd.i) if last else buf(c2, x.h // s_, x.w // s_)
                        new_block(r, c1, c2, k, s_, x, a)
                    x = a
                tens[nd.i] = x
            elif nd.type == "Bottleneck":
                c1, c2, *rest = nd.args
                shortcut = rest[0] if rest else True
                x = srcs[0]
                c_ = int(c2 * 0.5)
                for ri, r in enumerate(reps):
                    yb = out_of(nd.i) if ri == len(reps) - 1 else buf(c2, x.h, x.w)
                    t = buf(c_, x.h, x.w)
                    new_block(r + ".cv1", c1, c_, 1, 1, x, t)
                    new_block(r + ".cv2", c_, c2, 3, 1, t, yb, res=x if (shortcut and c1 == c2) else None)
                    x, c1 = yb, c2
                tens[nd.i] = x
            elif nd.type == "SPP":
                # models/common.py:281-290: cv2(cat[x, mp5(x), mp9(x), mp13(x)]) with x = cv1(input); each pool reads x
                c1, c2, *rest = nd.args
                ks = tuple(rest[0]) if rest else (5, 9, 13)
                x = srcs[0]
                c_ = c1 // 2
                cat = buf((len(ks) + 1) * c_, x.h, x.w)
                b1 = new_block(base + ".cv1", c1, c_, 1, 1, x, cat.slice(0, c_))
                for q, k in enumerate(ks):
                    idx = torch.zeros(n * x.h * x.w * c_, dtype=torch.uint8, device=dev)
                    self.keep.append(idx)
                    src, dst = cat.slice(0, c_), cat.slice((q + 1) * c_, c_)
                    b1.post_fwd.append(lambda src=src, dst=dst, k=k, idx=idx: T.maxpool_train_fwd(src, dst, k, idx))
                    b1.pre_bwd.append(lambda src=src, dst=dst, k=k, idx=idx: T.maxpool_bwd(self.grad_of(dst), self.grad_of(src),
                                                                                        k, idx, accumulate=True))
                y = out_of(nd.i)
                new_block(base + ".cv2", (len(ks) + 1) * c_, c2, 1, 1, cat, y)
                tens[nd.i] = y
            elif nd.type == "Upsample":
                tens[nd.i] = None
            elif nd.type == "Concat":
                tens[nd.i] = cat_buf[nd.i]

        # ---- Detect heads
        self.heads = []
        head_ld = ops.cout_pad(det.na * det.no)
        dec = _lib.DecodeDesc()
        for j, s in enumerate(nodes[-1].srcs):
            x = tens[s]
            hd = dict(x=x, c1=x.c, j=j)
            hd["out"] = torch.zeros(n * x.h * x.w, head_ld, dtype=torch.float32, device=dev)
            hd["raw"] = torch.zeros(n, det.na, x.h, x.w, det.no, dtype=torch.float32, device=dev)
            hd["wf"] = torch.zeros(head_ld, x.c, dtype=torch.bfloat16, device=dev)
            hd["wd"] = torch.zeros(ops.cout_pad(x.c), head_ld, dtype=torch.bfloat16, device=dev)
            hd["bias"] = torch.zeros(head_ld, dtype=torch.float32, device=dev)
            hd["zero_bi"] = torch.zeros(ops.cout_pad(x.c), dtype=torch.float32, device=dev)
            hd["dy"] = buf(head_ld, x.h, x.w)
            hd["dw"] = torch.zeros(head_ld, x.c, 1, 1, dtype=torch.float32, device=dev)
            self.heads.append(hd)
            lv = dec.levels[j]
            lv.head, lv.head_ld, lv.raw_out = hd["out"].data_ptr(), head_ld, hd["raw"].data_ptr()
            lv.ny, lv.nx, lv.stride = x.h, x.w, float(det.stride[j])
        dec.nl, dec.bs, dec.na, dec.no, dec.z = det.nl, n, det.na, det.no, None
        self.dec = dec
        self.err = torch.zeros(1, dtype=torch.int32, device=dev)

        # ---- static backward plan: for every activation tensor, is the first gradient contribution a write?
        self.param_names = []
        for b in self.blocks:
            self.param_names += [b.prefix + ".conv.weight", b.prefix + ".bn.weight", b.prefix + ".bn.bias"]
        for hd in self.heads:
            self.param_names += [f"model.{det.i}.m.{hd['j']}.weight", f"model.{det.i}.m.{hd['j']}.bias"]

    # ------------------------------------------------------------------------------------------------ helpers
    def _scratch(self, c, hh, ww, dev, tag=""):
        key = (c, hh, ww, tag)
        if key not in self.scratch:
            self.scratch[key] = PaddedNHWC.zeros(self.n, hh, ww, c, device=dev)
        return self.scratch[key]

    def grad_of(self, t: PaddedNHWC) -> PaddedNHWC:
        """Gradient buffer mirroring an activation buffer (same geometry, same channel slice)."""
        key = t.buf.data_ptr()
        g = self.grad_bufs.get(key)
        if g is None:
            g = self.grad_bufs[key] = PaddedNHWC(torch.zeros_like(t.buf), 0, t.buf.shape[3])
        return g.slice(t.coff, t.c)

    # ------------------------------------------------------------------------------------------------ CUDA graphs
    # Every launch of a step is stream-ordered with no host synchronisation, so after one eager (warm-up) step the whole
    # forward and the whole backward are each captured into a CUDA graph and replayed: ~700 launches per step cost two
    # graph launches on the host.  Master parameters, running statistics and all buffers keep their addresses.
    use_graphs = True

    def forward(self, x: torch.Tensor, in_div=0.0):
        if not self.use_graphs:
            return self._forward_impl(x, in_div)
        st = self.__dict__.setdefault("_gf", {"n": 0})
        if st["n"] == 0:
            st["n"] = 1
            return self._forward_impl(x, in_div)  # eager warm-up (function attributes, lazy allocations)
        if "graph" not in st:
            st["x"], st["div"] = x.clone(), in_div
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                st["out"] = self._forward_impl(st["x"], in_div)
            st["graph"] = g
        assert in_div == st["div"] and x.shape == st["x"].shape and x.dtype == st["x"].dtype
        st["x"].copy_(x)
        st["graph"].replay()
        return st["out"]

    def backward(self, graws):
        if not self.use_graphs:
            return self._backward_impl(graws)
        st = self.__dict__.setdefault("_gb", {"n": 0})
        if st["n"] == 0:
            st["n"] = 1
            return self._backward_impl(graws)
        if "graph" not in st:
            st["g"] = [g.detach().float().contiguous().clone() for g in graws]
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                st["out"] = self._backward_impl(st["g"])
            st["graph"] = gr
        for dst, src in zip(st["g"], graws):
            dst.copy_(src)
        st["graph"].replay()
        return [o.clone() for o in st["out"]]  # .grad must not alias buffers the next replay overwrites

    # ------------------------------------------------------------------------------------------------ forward
    def _forward_impl(self, x: torch.Tensor, in_div=0.0):
        P, det = self.P, self.model.detect
        T.im2col_first(x, self.im2col, in_div)
        for b in self.blocks:
            w = P[b.prefix + ".conv.weight"]
            if b.first:
                b.wf.zero_()
                b.wf[: b.c2, :27] = w.detach().reshape(b.c2, 27).to(torch.bfloat16)
                ops.conv_bn_act(b.x, b.wf, b.zero_b, b.c2, 1, 1, ops.ACT_NONE, out=b.y, err=self.err)
            else:
                T.pack_weights(w.detach(), b.wf, b.wd)
                ops.conv_bn_act(b.x, b.wf, b.zero_b, b.c2, b.k, b.s, ops.ACT_NONE, out=b.y, err=self.err)
            st = b.st
            T.bn_stats(b.y, st["sum"], st["sumsq"])
            count = self.n * b.y.h * b.y.w
            if self.sync_bn:  # nn.SyncBatchNorm (train.py:270-272): batch statistics over every rank's pixels
                dist.all_reduce(st["sums"])  # [sum | sumsq] share one buffer: one collective per layer
                count *= self.world
            T.bn_finalize(st["sum"], st["sumsq"], P[b.prefix + ".bn.weight"].detach(), P[b.prefix + ".bn.bias"].detach(),
                          count, st["scale"], st["shift"], st["mean"], st["rstd"],
                          P[b.prefix + ".bn.running_mean"], P[b.prefix + ".bn.running_var"])
            T.bn_act_fwd(b.y, st["scale"], st["shift"], b.a, b.res, b.upsample)
            for fn in b.post_fwd:
                fn()
        for hd in self.heads:
            w = P[f"model.{det.i}.m.{hd['j']}.weight"].detach()
            co = det.na * det.no
            hd["wf"][:co] = w.reshape(co, hd["c1"]).to(torch.bfloat16)
            hd["wd"][: hd["c1"], :co] = w.reshape(co, hd["c1"]).t().to(torch.bfloat16)
            hd["bias"][:co] = P[f"model.{det.i}.m.{hd['j']}.bias"].detach()
            ops.conv_bn_act(hd["x"], hd["wf"], hd["bias"], co, 1, 1, ops.ACT_NONE, out_f32=hd["out"], err=self.err)
        _lib.check(_lib.lib().y3_detect_head_decode_fwd(C.byref(self.dec), _stream()), "y3_detect_head_decode_fwd")
        return [hd["raw"] for hd in self.heads]

    # ------------------------------------------------------------------------------------------------ backward
    def _backward_impl(self, graws):
        """graws: dL/draw per level (fp32 [n,na,ny,nx,no]).  Returns gradients aligned with ``self.param_names``."""
        det = self.model.detect
        co = det.na * det.no
        written: set = set()

        def contribute_conv(dy, wd, zero_b, c_in, k, x):
            gx = self.grad_of(x)
            key = (x.buf.data_ptr(), x.coff, x.c)
            first = key not in written and not self._overlaps(written, key)
            ops.conv_bn_act(dy, wd, zero_b, c_in, k, 1, ops.ACT_NONE, out=gx, res=None if first else gx, err=self.err)
            written.add(key)

        def contribute_add(src, x):
            gx = self.grad_of(x)
            key = (x.buf.data_ptr(), x.coff, x.c)
            first = key not in written and not self._overlaps(written, key)
            T.add_nhwc(src, gx, accumulate=not first)
            written.add(key)

        head_grads = []
        for hd, g in zip(self.heads, graws):
            x = hd["x"]
            g = g.detach().float().contiguous()
            hd["dy"].buf[:, 1:-1, 1:-1, :co] = g.permute(0, 2, 3, 1, 4).reshape(self.n, x.h, x.w, co).to(torch.bfloat16)
            db = g.sum(dim=(0, 2, 3)).reshape(co)
            hd["dw"].zero_()
            T.conv_wgrad(hd["dy"], x, hd["dw"], 1)
            contribute_conv(hd["dy"], hd["wd"], hd["zero_bi"], hd["c1"], 1, x)
            head_grads.append((hd["dw"][:co].clone(), db))

        grads = {}
        for b in reversed(self.blocks):
            st = b.st
            for fn in b.pre_bwd:
                fn()
            da = self.grad_of(b.a)
            if self.sync_bn:
                # local sums are the (rank-local) gamma/beta gradients; dy needs the sums over all ranks
                T.bn_act_bwd(b.y, da, b.dy, st["scale"], st["shift"], st["mean"], st["rstd"], st["dbeta"], st["dgamma"],
                             b.upsample, phase=1)
                st["gsums"].copy_(st["dsums"])
                dist.all_reduce(st["gsums"])
                T.bn_act_bwd(b.y, da, b.dy, st["scale"], st["shift"], st["mean"], st["rstd"], st["gsums"][0], st["gsums"][1],
                             b.upsample, phase=2, count=self.n * b.y.h * b.y.w * self.world)
            else:
                T.bn_act_bwd(b.y, da, b.dy, st["scale"], st["shift"], st["mean"], st["rstd"], st["dbeta"], st["dgamma"],
                             b.upsample)
            b.dw.zero_()
            src = b.dy
            if b.s == 2:
                src = T.zero_stuff(b.dy, b.dy_up)
            kk = 1 if b.first else b.k
            T.conv_wgrad(src, b.x, b.dw, kk, tap_major=b.dw_tm)
            dw = b.dw.permute(1, 2, 0).reshape(b.c2, -1, kk, kk) if b.dw_tm else b.dw  # -> [co, ci, k, k]
            if b.first:
                grads[b.prefix + ".conv.weight"] = dw[:, :27].reshape(b.c2, 3, 3, 3).clone()
            else:
                grads[b.prefix + ".conv.weight"] = dw.contiguous().clone() if b.dw_tm and kk > 1 else dw.clone()
                contribute_conv(src, b.wd, b.zero_bi, b.c1, b.k, b.x)
            grads[b.prefix + ".bn.weight"] = st["dgamma"].clone()
            grads[b.prefix + ".bn.bias"] = st["dbeta"].clone()
            if b.res is not None:
                contribute_add(da, b.res)  # Bottleneck shortcut: the block output's gradient also flows to its input
        for hd, (dwh, dbh) in zip(self.heads, head_grads):
            grads[f"model.{det.i}.m.{hd['j']}.weight"] = dwh
            grads[f"model.{det.i}.m.{hd['j']}.bias"] = dbh
        return [grads[k] for k in self.param_names]

    @staticmethod
    def _overlaps(written, key):
        ptr, coff, c = key
        return any(p == ptr and not (coff + c <= o or o + cc <= coff) for (p, o, cc) in written)

    def check_errors(self):
        e = int(self.err.item())
        if e:
            raise _lib.Y3Error(f"device watchdog reported pipeline stall code {e}")


class TrainFn(torch.autograd.Function):
    """pred = model(imgs) in train mode as ONE autograd node: backward() runs TrainEngine.backward."""

    @staticmethod
    def forward(ctx, engine, x, in_div, *params):
        ctx.engine = engine
        raws = engine.forward(x, in_div)
        return tuple(r.clone() for r in raws)

    @staticmethod
    def backward(ctx, *graws):
        grads = ctx.engine.backward(graws)
        return (None, None, None, *grads)
