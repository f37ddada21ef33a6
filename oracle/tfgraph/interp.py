"""numpy interpreter for TensorFlow-1.3 GraphDefs  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Purpose: execute the reference's OWN serialized graph (``/root/reference/model/air-model.meta``, written by
``training.py:141`` through ``tf.train.Saver``; it holds the complete training graph ``air/...`` including the autodiff
gradient graph, global-norm clipping and the 36 ``ApplyAdam`` ops, and the test-mode graph ``air_1/...``) without
TensorFlow.  Everything structural -- op order, constants, wiring of the while loop, which gradient function TF
generated for every op, the stop masks -- therefore comes from the reference artifact itself and not from a re-reading
of its Python source.  What IS restated here is the arithmetic of each TF kernel (one numpy fp32 expression per op,
rounded op by op like TF's executor; the per-kernel semantics follow tensorflow/core/kernels of TF 1.3).

Execution model: demand driven (pull) with memoisation per (node, loop iteration).  ``tf.while_loop`` frames
(Enter / Merge / Switch / LoopCond / NextIteration / Exit) are evaluated iteration by iteration: a Merge at iteration
0 takes its Enter input, at iteration i its NextIteration input evaluated at i-1; an Exit takes the Merge value of the
first iteration whose LoopCond is false.  Control inputs are honoured (evaluated before the op), which is what makes
the forward loop push every stashed activation (``StackPush`` hangs off the forward counter by control edges) before
the gradient loop pops it.  Frames are not nested in this graph (asserted).

Random ops, queue outputs, placeholders and variables are supplied by the caller.
"""
from __future__ import annotations

import math
import sys
import threading

import numpy as np

from . import pb

f32 = np.float32


def _strided_index(shape, begin, end, strides, a):
    """StridedSlice masks -> (python index tuple, final_shape) following tensorflow/core/util/strided_slice_op.cc."""
    bm, em, elm, nam, sam = (int(a.get(k, 0)) for k in ("begin_mask", "end_mask", "ellipsis_mask", "new_axis_mask",
                                                         "shrink_axis_mask"))
    n = len(begin)
    idx = []
    dim = 0
    n_real = sum(1 for i in range(n) if not ((nam >> i) & 1) and not ((elm >> i) & 1))
    for i in range(n):
        if (elm >> i) & 1:
            fill = len(shape) - n_real
            idx += [slice(None)] * fill
            dim += fill
        elif (nam >> i) & 1:
            idx.append(None)
        elif (sam >> i) & 1:
            b = int(begin[i])
            idx.append(b if b >= 0 else b + shape[dim])
            dim += 1
        else:
            s = int(strides[i])
            b = None if (bm >> i) & 1 else int(begin[i])
            e = None if (em >> i) & 1 else int(end[i])
            idx.append(slice(b, e, s))
            dim += 1
    return tuple(idx)


def _reduce(fn, x, axes, keep):
    axes = tuple(int(v) % max(x.ndim, 1) for v in np.atleast_1d(axes)) if x.ndim else ()
    if not axes:
        return x.copy() if not keep or x.ndim == 0 else x.copy()
    return fn(x, axis=axes, keepdims=keep)


def _softplus(x):
    thr = x.dtype.type(math.log(np.finfo(x.dtype).eps) + 2.0)
    e = np.exp(x)
    mid = np.log(e + x.dtype.type(1.0))
    return np.where(x > -thr, x, np.where(x < thr, e, mid)).astype(x.dtype)


class Interpreter:
    def __init__(self, path_or_nodes, variables=None, feeds=None, random_fn=None, float_dtype=np.float32):
        """``float_dtype=np.float64`` evaluates the same graph with every float32 constant, variable, feed, noise
        tensor and Cast target widened to fp64 (the "truth" run; TF itself would run it in fp32)."""
        self.fdt = np.dtype(float_dtype).type
        if isinstance(path_or_nodes, str):
            self.nodes, self.version = pb.load_metagraph(path_or_nodes)
        else:
            self.nodes, self.version = path_or_nodes
        self.variables = {k: self._widen(v) for k, v in (variables or {}).items()}   # VariableV2 name -> ndarray
        self.feeds = {k: self._widen(v) for k, v in (feeds or {}).items()}           # "node:idx" -> ndarray
        self.random_fn = random_fn                      # (node, iteration, shape) -> ndarray
        self.memo = {}
        self.frame_memo = {}
        self.loop_cond = {}                             # frame -> LoopCond node name
        self.trip = {}
        self.stacks = {}
        self.stack_total = {}
        self.assigned = {}                              # variable name -> new value (ApplyAdam / Assign / AssignAdd)
        for nd in self.nodes.values():
            if nd.op == "Switch" and self.nodes[nd.inputs[1][0]].op == "LoopCond":
                self.loop_cond[self.frame(nd.name)] = nd.inputs[1][0]

    def _widen(self, v):
        v = np.asarray(v)
        return v.astype(self.fdt) if v.dtype == np.float32 and self.fdt is not np.float32 else v

    def _const(self, v):
        """fp64 mode: a float32 constant becomes the shortest decimal that round-trips to it, i.e. the Python literal
        the reference wrote (0.05, 1e-9, 0.99 ...), not the float32 rounding of it."""
        if v.dtype != np.float32 or self.fdt is np.float32:
            return v
        return np.array([float(np.format_float_positional(x, unique=True)) if np.isfinite(x) else float(x)
                         for x in v.ravel()], np.float64).reshape(v.shape)

    # ---- frames ---------------------------------------------------------------------------------------------
    def frame(self, name):
        """While-loop frame (Enter's frame_name) a node executes in; None = the outer graph."""
        stack = [name]
        while stack:
            cur = stack[-1]
            if cur in self.frame_memo:
                stack.pop()
                continue
            nd = self.nodes[cur]
            if nd.op in ("Enter", "RefEnter"):
                self.frame_memo[cur] = nd.attr["frame_name"].decode()
                stack.pop()
                continue
            if nd.op in ("Exit", "RefExit"):
                self.frame_memo[cur] = None
                stack.pop()
                continue
            if nd.op == "Merge":                        # loop Merge: the Enter side decides (breaks the cycle)
                deps = [i for i, _ in nd.inputs if self.nodes[i].op != "NextIteration"][:1]
            else:
                deps = [i for i, _ in nd.inputs][:1] or nd.ctrl[:1]
            if not deps:
                self.frame_memo[cur] = None
                stack.pop()
                continue
            d = deps[0]
            if d in self.frame_memo:
                self.frame_memo[cur] = self.frame_memo[d]
                stack.pop()
            else:
                stack.append(d)
        return self.frame_memo[name]

    def trip_count(self, fr):
        if fr not in self.trip:
            lc = self.loop_cond[fr]
            i = 0
            while bool(self.eval(lc, 0, i)):
                i += 1
                assert i < 10000
            self.trip[fr] = i
        return self.trip[fr]

    # ---- evaluation -----------------------------------------------------------------------------------------
    def fetch(self, tensor):
        name, _, idx = tensor.partition(":")
        assert self.frame(name) is None, f"{name} lives inside a loop frame"
        return self.eval(name, int(idx or 0), None)

    def eval(self, name, idx, it):
        key = f"{name}:{idx}"
        if key in self.feeds:
            return self.feeds[key]
        mk = (name, it)
        if mk not in self.memo:
            self.memo[mk] = self._run(self.nodes[name], it)
        out = self.memo[mk]
        return out[idx] if isinstance(out, tuple) else out

    def _in(self, nd, k, it):
        name, idx = nd.inputs[k]
        return self.eval(name, idx, it)

    def _run(self, nd, it):
        op = nd.op
        if op == "Const":
            return self._const(nd.attr["value"])
        if op in ("Enter", "RefEnter"):
            assert self.frame(nd.inputs[0][0]) is None, "nested while frames are not supported"
            return self._in(nd, 0, None)
        if op == "Merge":
            ni = [k for k, (i, _) in enumerate(nd.inputs) if self.nodes[i].op == "NextIteration"]
            assert len(ni) == 1 and len(nd.inputs) == 2, f"non-loop Merge {nd.name}"
            return self._in(nd, 1 - ni[0], 0) if it == 0 else self._in(nd, ni[0], it - 1)
        if op == "Exit":
            sw = self.nodes[nd.inputs[0][0]]
            fr = self.frame(sw.name)
            n = self.trip_count(fr)
            return self._in(sw, 0, n)
        # data inputs first, control inputs second: the forward counter's Add carries the StackPush ops of ITS
        # iteration as control inputs and reaches the previous iteration's pushes through its data input
        args = [self._in(nd, k, it) for k in range(len(nd.inputs))]
        for c in nd.ctrl:
            self.eval(c, 0, it if self.frame(c) is not None else None)
        if op == "Switch":
            return (args[0], args[0])                   # consumers are reached only on the live side
        if op in ("Identity", "NextIteration", "LoopCond", "StopGradient", "PreventGradient"):
            return args[0]
        if op in ("NoOp", "ControlTrigger"):
            return None
        fn = getattr(self, "op_" + op, None)
        if fn is None:
            raise NotImplementedError(f"op {op} ({nd.name})")
        return fn(nd, it, *args)

    # ---- sources --------------------------------------------------------------------------------------------
    def op_VariableV2(self, nd, it):
        return self.variables[nd.name]

    def op_Placeholder(self, nd, it):
        raise KeyError(f"placeholder {nd.name} not fed")

    def op_QueueDequeueManyV2(self, nd, it, *a):
        raise KeyError(f"queue output {nd.name} not fed")

    def op_RandomShuffleQueueV2(self, nd, it):
        return None

    def op_RandomStandardNormal(self, nd, it, shape):
        return np.asarray(self.random_fn(nd.name, it, tuple(int(v) for v in shape))).astype(self.fdt)

    op_RandomUniform = op_RandomStandardNormal

    # ---- stacks / tensor arrays -----------------------------------------------------------------------------
    def op_Stack(self, nd, it):
        self.stacks[nd.name] = []
        return self.stacks[nd.name]

    def op_StackPush(self, nd, it, h, v):
        assert len(h) == it, f"{nd.name}: push of iteration {it} onto a stack of {len(h)}"
        h.append(v)
        return v

    def op_StackPop(self, nd, it, h):
        total = self.stack_total.setdefault(id(h), len(h))     # pushes are complete when the first pop runs
        assert len(h) + it == total, f"{nd.name}: pops out of order"
        return h.pop()

    def op_TensorArrayV3(self, nd, it, size):
        ta = {}
        return (ta, f32(0.0))

    def op_TensorArrayWriteV3(self, nd, it, h, index, value, flow):
        h[int(index)] = value
        return f32(0.0)

    def op_TensorArraySizeV3(self, nd, it, h, flow):
        return np.int32(len(h))

    def op_TensorArrayGatherV3(self, nd, it, h, indices, flow):
        return np.stack([h[int(i)] for i in indices]) if len(indices) else np.zeros((0,), f32)

    # ---- shapes ---------------------------------------------------------------------------------------------
    def op_Shape(self, nd, it, x):
        return np.array(np.shape(x), np.int32)

    def op_ShapeN(self, nd, it, *xs):
        return tuple(np.array(np.shape(x), np.int32) for x in xs)

    def op_Size(self, nd, it, x):
        return np.int32(np.size(x))

    def op_Rank(self, nd, it, x):
        return np.int32(np.ndim(x))

    def op_Reshape(self, nd, it, x, shape):
        return np.reshape(x, tuple(int(v) for v in shape))

    def op_ExpandDims(self, nd, it, x, axis):
        return np.expand_dims(x, int(axis))

    def op_Squeeze(self, nd, it, x):
        dims = nd.attr.get("squeeze_dims") or None
        return np.squeeze(x, tuple(int(d) for d in dims) if dims else None)

    def op_Transpose(self, nd, it, x, perm):
        return np.ascontiguousarray(np.transpose(x, tuple(int(p) for p in perm)))

    def op_Pack(self, nd, it, *xs):
        return np.stack(xs, int(nd.attr.get("axis", 0)))

    def op_Unpack(self, nd, it, x):
        ax = int(nd.attr.get("axis", 0))
        return tuple(np.take(x, i, ax) for i in range(x.shape[ax]))

    def op_ConcatV2(self, nd, it, *xs):
        return np.concatenate(xs[:-1], int(xs[-1]))

    def op_ConcatOffset(self, nd, it, dim, *shapes):
        out, off = [], 0
        for s in shapes:
            o = np.zeros_like(s)
            o[int(dim)] = off
            off += int(s[int(dim)])
            out.append(o)
        return tuple(out)

    def op_Split(self, nd, it, dim, x):
        return tuple(np.split(x, int(nd.attr["num_split"]), int(dim)))

    def op_Slice(self, nd, it, x, begin, size):
        idx = tuple(slice(int(b), None if int(s) == -1 else int(b) + int(s)) for b, s in zip(begin, size))
        return x[idx]

    def op_StridedSlice(self, nd, it, x, begin, end, strides):
        return np.asarray(x)[_strided_index(np.shape(x), begin, end, strides, nd.attr)]

    def op_StridedSliceGrad(self, nd, it, shape, begin, end, strides, dy):
        out = np.zeros(tuple(int(v) for v in shape), dy.dtype)
        idx = _strided_index(out.shape, begin, end, strides, nd.attr)
        out[idx] = np.reshape(dy, out[idx].shape)
        return out

    def op_Tile(self, nd, it, x, mult):
        return np.tile(x, tuple(int(m) for m in mult))

    def op_Pad(self, nd, it, x, pads):
        return np.pad(x, [(int(a), int(b)) for a, b in pads])

    def op_Fill(self, nd, it, dims, v):
        return np.full(tuple(int(d) for d in dims), v, np.asarray(v).dtype)

    def op_ZerosLike(self, nd, it, x):
        return np.zeros_like(x)

    def op_Range(self, nd, it, s, l, d):
        return np.arange(s, l, d, dtype=np.asarray(s).dtype)

    def op_LinSpace(self, nd, it, start, stop, num):
        n, t = int(num), self.fdt
        start, stop = t(start), t(stop)
        if n == 1:
            return np.array([start], t)
        step = t((stop - start) / t(n - 1))
        return (start + step * np.arange(n, dtype=t)).astype(t)

    def op_Cast(self, nd, it, x):
        dst = pb.DT[nd.attr["DstT"][1]]
        return np.asarray(x).astype(self.fdt if dst is np.float32 else dst)

    def op_BroadcastGradientArgs(self, nd, it, s0, s1):
        n = max(len(s0), len(s1))
        a = [1] * (n - len(s0)) + [int(v) for v in s0]
        b = [1] * (n - len(s1)) + [int(v) for v in s1]
        # every size-1 (or missing) dimension is reduced, also where both sides are 1: harmless, the result is
        # reshaped to the operand's shape by the Reshape that always follows in a gradient graph
        r0 = [i for i in range(n) if a[i] == 1]
        r1 = [i for i in range(n) if b[i] == 1]
        return (np.array(r0, np.int32), np.array(r1, np.int32))

    # ---- elementwise ----------------------------------------------------------------------------------------
    def op_Add(self, nd, it, a, b):
        return a + b

    def op_Sub(self, nd, it, a, b):
        return a - b

    def op_Mul(self, nd, it, a, b):
        return a * b

    def op_RealDiv(self, nd, it, a, b):
        return a / b

    def op_FloorDiv(self, nd, it, a, b):
        return np.floor_divide(a, b)

    def op_FloorMod(self, nd, it, a, b):
        return np.mod(a, b)

    def op_Pow(self, nd, it, a, b):
        return np.power(a, b)

    def op_Maximum(self, nd, it, a, b):
        return np.maximum(a, b)

    def op_Minimum(self, nd, it, a, b):
        return np.minimum(a, b)

    def op_Neg(self, nd, it, a):
        return -a

    def op_Floor(self, nd, it, a):
        return np.floor(a)

    def op_Round(self, nd, it, a):
        return np.rint(a)                               # half to even, like TF >= 1.0

    def op_Log(self, nd, it, a):
        return np.log(a)

    def op_Exp(self, nd, it, a):
        return np.exp(a)

    def op_Sqrt(self, nd, it, a):
        return np.sqrt(a)

    def op_Square(self, nd, it, a):
        return a * a

    def op_Reciprocal(self, nd, it, a):
        return (f32(1.0) / a) if a.dtype == np.float32 else 1 / a

    def op_Sigmoid(self, nd, it, a):
        return (f32(1.0) / (f32(1.0) + np.exp(-a))).astype(a.dtype)

    def op_Tanh(self, nd, it, a):
        return np.tanh(a)

    def op_Relu(self, nd, it, a):
        return np.maximum(a, a.dtype.type(0))

    def op_Softplus(self, nd, it, a):
        return _softplus(a)

    def op_SigmoidGrad(self, nd, it, y, dy):
        return dy * y * (f32(1.0) - y)

    def op_TanhGrad(self, nd, it, y, dy):
        return dy * (f32(1.0) - y * y)

    def op_SqrtGrad(self, nd, it, y, dy):
        return dy * f32(0.5) / y

    def op_ReluGrad(self, nd, it, g, x):
        return np.where(x > 0, g, f32(0.0)).astype(g.dtype)

    def op_SoftplusGrad(self, nd, it, g, x):
        return g / (np.exp(-x) + f32(1.0))

    def op_Less(self, nd, it, a, b):
        return a < b

    def op_LessEqual(self, nd, it, a, b):
        return a <= b

    def op_Greater(self, nd, it, a, b):
        return a > b

    def op_GreaterEqual(self, nd, it, a, b):
        return a >= b

    def op_Equal(self, nd, it, a, b):
        return a == b

    def op_LogicalAnd(self, nd, it, a, b):
        return np.logical_and(a, b)

    def op_LogicalNot(self, nd, it, a):
        return np.logical_not(a)

    def op_Select(self, nd, it, c, t, e):
        c = np.asarray(c)
        if c.ndim == 1 and np.ndim(t) > 1:              # tf.where with a vector condition selects rows
            c = c.reshape((-1,) + (1,) * (np.ndim(t) - 1))
        return np.where(c, t, e).astype(np.asarray(t).dtype)

    def op_AddN(self, nd, it, *xs):
        acc = xs[0]
        for x in xs[1:]:
            acc = acc + x
        return acc

    # ---- reductions / contractions --------------------------------------------------------------------------
    def op_Sum(self, nd, it, x, axes):
        return _reduce(np.sum, np.asarray(x), axes, bool(nd.attr.get("keep_dims", False)))

    def op_Prod(self, nd, it, x, axes):
        return _reduce(np.prod, np.asarray(x), axes, bool(nd.attr.get("keep_dims", False)))

    def op_Mean(self, nd, it, x, axes):
        x = np.asarray(x)
        return _reduce(np.mean, x, axes, bool(nd.attr.get("keep_dims", False))).astype(x.dtype)

    def op_Any(self, nd, it, x, axes):
        return _reduce(np.any, np.asarray(x), axes, bool(nd.attr.get("keep_dims", False)))

    def op_L2Loss(self, nd, it, x):
        return (np.sum(x * x) / f32(2.0)).astype(x.dtype)

    def op_MatMul(self, nd, it, a, b):
        if nd.attr.get("transpose_a"):
            a = a.T
        if nd.attr.get("transpose_b"):
            b = b.T
        return np.matmul(a, b)

    def op_BatchMatMul(self, nd, it, a, b):
        if nd.attr.get("adj_x"):
            a = np.swapaxes(a, -1, -2)
        if nd.attr.get("adj_y"):
            b = np.swapaxes(b, -1, -2)
        # products rounded, then summed in k order (no FMA): einsum over k of explicit products
        prod = a[..., :, :, None] * b[..., None, :, :]
        acc = prod[..., :, 0, :]
        for k in range(1, prod.shape[-2]):
            acc = acc + prod[..., :, k, :]
        return acc

    def op_BiasAdd(self, nd, it, x, b):
        return x + b

    def op_BiasAddGrad(self, nd, it, g):
        return np.sum(g.reshape(-1, g.shape[-1]), axis=0)

    def op_Gather(self, nd, it, params, indices):
        return np.take(params, indices, axis=0)

    def op_ResizeBilinear(self, nd, it, x, size):
        """resize_bilinear_op.cc (TF 1.3): in = out_index * (in_size / out_size) (align_corners false), lower =
        floor(in), upper = min(lower + 1, in_size - 1), lerp = in - lower; rows are interpolated along x first
        (top, bottom), then along y, each as a + (b - a) * lerp."""
        assert not nd.attr.get("align_corners")
        oh, ow = int(size[0]), int(size[1])
        B, H, W, C = x.shape

        def axis(n_in, n_out):
            scale = f32(n_in) / f32(n_out)
            src = np.arange(n_out, dtype=f32) * scale
            lo = np.floor(src).astype(np.int64)
            return lo, np.minimum(lo + 1, n_in - 1), (src - lo.astype(f32)).astype(x.dtype)
        y0, y1, ly = axis(H, oh)
        x0, x1, lx = axis(W, ow)
        lx = lx[None, None, :, None]
        ly = ly[None, :, None, None]
        tl, tr = x[:, y0][:, :, x0], x[:, y0][:, :, x1]
        bl, br = x[:, y1][:, :, x0], x[:, y1][:, :, x1]
        top = tl + (tr - tl) * lx
        bottom = bl + (br - bl) * lx
        return top + (bottom - top) * ly

    def op_DrawBoundingBoxes(self, nd, it, images, boxes):
        """draw_bounding_box_op.cc (TF 1.3): box (ymin, xmin, ymax, xmax) in [0,1] -> rows/cols * (size - 1); the
        four border lines are painted with colour table entry (box index % 10); the first entry is yellow
        (1, 1, 0, 1), of which a 1-channel image takes the first component."""
        table = [(1, 1, 0, 1), (0, 0, 1, 1), (1, 0, 0, 1), (0, 1, 0, 1), (0.5, 0, 0.5, 1), (0.5, 0.5, 0, 1),
                 (0.5, 0, 0, 1), (0, 0, 0.5, 1), (0, 1, 1, 1), (1, 0, 1, 1)]
        out = images.copy()
        B, H, W, C = out.shape
        for b in range(B):
            for bb in range(boxes.shape[1]):
                col = np.array(table[bb % 10][:C], out.dtype)
                ymin, xmin, ymax, xmax = (float(v) for v in boxes[b, bb])
                r0, r1 = int(ymin * (H - 1)), int(ymax * (H - 1))
                c0, c1 = int(xmin * (W - 1)), int(xmax * (W - 1))
                if r0 > r1 or c0 > c1 or r0 >= H or r1 < 0 or c0 >= W or c1 < 0:
                    continue
                rr0, rr1, cc0, cc1 = max(r0, 0), min(r1, H - 1), max(c0, 0), min(c1, W - 1)
                if r0 >= 0:
                    out[b, r0, cc0:cc1 + 1] = col
                if r1 < H:
                    out[b, r1, cc0:cc1 + 1] = col
                if c0 >= 0:
                    out[b, rr0:rr1 + 1, c0] = col
                if c1 < W:
                    out[b, rr0:rr1 + 1, c1] = col
        return out

    def op_Where(self, nd, it, c):
        return np.argwhere(c).astype(np.int64)

    def op_UnsortedSegmentSum(self, nd, it, data, ids, num):
        out = np.zeros((int(num),) + data.shape[ids.ndim:], data.dtype)
        np.add.at(out, ids, data)                       # unbuffered, sequential in index order
        return out

    def op_DynamicStitch(self, nd, it, *a):
        n = len(a) // 2
        idxs, data = a[:n], a[n:]
        size = max(int(np.max(i)) for i in idxs if np.size(i)) + 1
        first = next(d for d, i in zip(data, idxs) if np.size(i))
        out = np.zeros((size,) + first.shape[np.ndim(idxs[0]):], first.dtype)
        for i, d in zip(idxs, data):
            out[np.asarray(i)] = d
        return out

    # ---- state ----------------------------------------------------------------------------------------------
    def op_ApplyAdam(self, nd, it, var, m, v, b1p, b2p, lr, b1, b2, eps, g):
        """training_ops.cc ApplyAdam (non-Nesterov): alpha = lr*sqrt(1-b2p)/(1-b1p); m += (g-m)(1-b1);
        v += (g*g-v)(1-b2); var -= m*alpha/(sqrt(v)+eps)."""
        one = self.fdt(1.0)
        alpha = self.fdt(lr * np.sqrt(one - b2p) / (one - b1p))
        m2 = m + (g - m) * (one - b1)
        v2 = v + (g * g - v) * (one - b2)
        var2 = var - (m2 * alpha) / (np.sqrt(v2) + eps)
        names = [i for i, _ in nd.inputs[:3]]
        for n_, val in zip(names, (var2, m2, v2)):
            self.assigned[n_] = val
        return var2

    def op_Assign(self, nd, it, ref, val):
        self.assigned[nd.inputs[0][0]] = val
        return val

    def op_AssignAdd(self, nd, it, ref, val):
        out = ref + val
        self.assigned[nd.inputs[0][0]] = out
        return out


def run_in_big_stack(fn, *a, **k):
    """Deep graphs recurse deeply: run ``fn`` in a thread with a 1 GB stack and a high recursion limit."""
    res = {}

    def target():
        try:
            res["v"] = fn(*a, **k)
        except BaseException as e:                      # noqa: BLE001 - re-raised in the caller
            res["e"] = e

    old = sys.getrecursionlimit()
    sys.setrecursionlimit(1_000_000)
    threading.stack_size(1 << 30)
    t = threading.Thread(target=target)
    t.start()
    t.join()
    threading.stack_size(0)
    sys.setrecursionlimit(old)
    if "e" in res:
        raise res["e"]
    return res["v"]
