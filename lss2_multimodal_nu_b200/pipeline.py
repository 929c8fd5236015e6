"""LiftSplatStep: the fixed-shape hot loop as a pre-allocated, CUDA-graph-captured step.

Training and benchmarking call the same shapes every iteration, so everything that can be
fixed is fixed once: device buffers, the workspace, and ONE CUDA graph holding the C-ABI call
sequence of a whole forward + backward of the path

      +-- lss_feat_stage (side stream) --------+
      |                                        v
  in -+-- lss_build_plan (K0, K1', sort, K3) ----+--> lss_liftsplat_fwd --> lss_liftsplat_bwd

(the staging copies depend only on the features and the plan only on the calibration, so the
two branches run concurrently inside the graph).  ``run()`` replays the graph on device-resident
inputs; ``HostPipeline`` feeds it from pinned HOST buffers with one packed copy per direction, both
inside the slot's graph, and keeps several steps in flight, so copies overlap kernels.

Nothing here changes results: it is the sequence functional.build_plan + functional.lift_splat
(autograd forward and backward) issue, minus the per-call allocations and Python overhead.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _abi
from .functional import GridSpec

CAL = ("rots", "trans", "intrins", "post_rots", "post_trans")


def _cal_sizes(B: int, N: int) -> Dict[str, int]:
    return {"rots": B * N * 9, "trans": B * N * 3, "intrins": B * N * 9, "post_rots": B * N * 9,
            "post_trans": B * N * 3}


class LiftSplatStep:
    """One forward + backward of lift+splat for fixed (B, N, D, fH, fW, C, grid)."""

    def __init__(self, B: int, N: int, D: int, fH: int, fW: int, C: int, grid: GridSpec,
                 us: torch.Tensor, vs: torch.Tensor, ds: torch.Tensor, device=None,
                 capture: bool = True, stream: Optional[torch.cuda.Stream] = None):
        dev = torch.device(device if device is not None else us.device)
        if dev.type != "cuda":
            raise RuntimeError("LiftSplatStep runs on CUDA only (no CPU fallback)")
        self.dev, self.grid = dev, grid
        self.B, self.N, self.D, self.fH, self.fW, self.C = B, N, D, fH, fW, C
        # every buffer below is created (and, where needed, zero-filled) on the step's own stream:
        # the warm run, the capture and load() all use that stream, so they are ordered behind the fills
        self._stream = stream if stream is not None else torch.cuda.Stream(dev)
        self._stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.device(dev), torch.cuda.stream(self._stream):
            self._allocate(us, vs, ds)
        self._side = torch.cuda.Stream(dev)
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._graph_cached: Optional[torch.cuda.CUDAGraph] = None
        self.kernels_per_step = 4 + 1 + 1 + 1      # plan (cells, scan, scatter, order), feature staging, fwd, bwd
        # warm run outside capture (module load, function attributes), then capture
        with torch.cuda.device(dev):
            with torch.cuda.stream(self._stream):
                self._enqueue(self._stream, self._side)
            self._stream.synchronize()
            if capture:
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph, stream=self._stream):
                    self._enqueue(torch.cuda.current_stream(dev), self._side)

    def _allocate(self, us, vs, ds) -> None:
        dev, grid = self.dev, self.grid
        B, N, D, fH, fW, C = self.B, self.N, self.D, self.fH, self.fW, self.C
        self.P = B * N * D * fH * fW
        X, Y, Z = grid.nx
        BN, HW = B * N, fH * fW
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        self.us, self.vs, self.ds = (t.to(dev).float().contiguous() for t in (us, vs, ds))
        # one packed input block: [calibration | depth | feat]  -> a single H2D copy per step
        sizes = _cal_sizes(B, N)
        self._layout, off = {}, 0
        for k in CAL:
            self._layout[k] = (off, sizes[k]); off += sizes[k]
        off = (off + 3) // 4 * 4                                   # keep depth/feat 16-byte aligned
        self._layout["depth"] = (off, BN * D * HW); off += BN * D * HW
        off = (off + 3) // 4 * 4
        self._layout["feat"] = (off, BN * C * HW); off += BN * C * HW
        self.in_block = torch.zeros(off, **f32)
        shp = {"rots": (B, N, 3, 3), "trans": (B, N, 3), "intrins": (B, N, 3, 3),
               "post_rots": (B, N, 3, 3), "post_trans": (B, N, 3), "depth": (BN, D, fH, fW),
               "feat": (BN, C, fH, fW)}
        self.inputs = {k: self.in_block[o:o + n].view(shp[k]) for k, (o, n) in self._layout.items()}
        self.in_shapes = shp
        # packed output block: [d_depth | d_feat]
        self.out_block = torch.empty(BN * D * HW + BN * C * HW, **f32)
        self.ddepth = self.out_block[:BN * D * HW].view(BN, D, fH, fW)
        self.dfeat = self.out_block[BN * D * HW:].view(BN, C, fH, fW)
        # upstream gradient and BEV map, channels-innermost storage of the logical (B, C*Z, X, Y)
        self._dbev = torch.zeros((B, X, Y, Z * C), **f32)
        self._bev = torch.empty((B, X, Y, Z * C), **f32)
        # plan + staging buffers
        self.cells = torch.empty(self.P, **i32)
        self.sorted_rec = torch.empty((self.P, 2), **i32)
        self.key_start = torch.empty(int(_abi.load().lss_plan_key_count(grid.c(), B)) + 1, **i32)
        self.counts = torch.zeros(2, **i32)
        self.feat_t = torch.empty((BN * HW, C), **f32)
        self._shape = _abi.make_shape(B, N, D, fH, fW, C)
        self._g = grid.c()
        nbytes = _abi.load().lss_plan_workspace_bytes(self._shape, self._g)
        if nbytes == 0:
            raise RuntimeError("lss_plan_workspace_bytes rejected the shape/grid")
        self._ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)

    # ---- views ---------------------------------------------------------------------------
    @property
    def bev(self) -> torch.Tensor:
        """Logical (B, C*Z, X, Y) result (channels_last strides) of the last run."""
        return self._bev.permute(0, 3, 1, 2)

    @property
    def dbev(self) -> torch.Tensor:
        """Logical (B, C*Z, X, Y) upstream-gradient buffer; write into it (e.g. copy_) before run()."""
        return self._dbev.permute(0, 3, 1, 2)

    @property
    def stream(self) -> torch.cuda.Stream:
        return self._stream

    # ---- the C-ABI call sequence ---------------------------------------------------------
    def enqueue_stage(self, st: int) -> None:
        p = lambda t: t.data_ptr()
        _abi.call("lss_feat_stage", p(self.inputs["feat"]), self.C * self.fH * self.fW, self._shape, _abi.LSS_F32,
                  p(self.feat_t), st)

    def enqueue_plan(self, st: int) -> None:
        p = lambda t: t.data_ptr()
        i = self.inputs
        _abi.call("lss_build_plan", p(self.us), p(self.vs), p(self.ds), p(i["rots"]), p(i["trans"]),
                  p(i["intrins"]), p(i["post_rots"]), p(i["post_trans"]), self._g, self._shape,
                  p(self.cells), p(self.key_start), p(self.sorted_rec), p(self.counts), p(self._ws),
                  self._ws.numel(), st)

    def enqueue_fwd(self, st: int) -> None:
        p = lambda t: t.data_ptr()
        _abi.call("lss_liftsplat_fwd", p(self.inputs["depth"]), self.D * self.fH * self.fW, _abi.LSS_F32,
                  p(self.feat_t), p(self.sorted_rec), p(self.key_start), self._g, self._shape, p(self._bev), st)

    def enqueue_bwd(self, st: int) -> None:
        p = lambda t: t.data_ptr()
        HW = self.fH * self.fW
        _abi.call("lss_liftsplat_bwd", p(self._dbev), p(self.inputs["depth"]), self.D * HW, _abi.LSS_F32,
                  p(self.feat_t), p(self.cells), self._g, self._shape, 0, _abi.LSS_F32, p(self.ddepth), self.D * HW,
                  p(self.dfeat), self.C * HW, st)

    def _enqueue(self, main: torch.cuda.Stream, side: torch.cuda.Stream, with_plan: bool = True) -> None:
        if with_plan:
            side.wait_stream(main)                  # fork: staging depends only on the features
            self.enqueue_stage(side.cuda_stream)
            self.enqueue_plan(main.cuda_stream)     # the plan depends only on the calibration
            main.wait_stream(side)                  # join
        else:
            self.enqueue_stage(main.cuda_stream)
        self.enqueue_fwd(main.cuda_stream)
        self.enqueue_bwd(main.cuda_stream)

    def run(self, wait_current: bool = False) -> None:
        """Enqueue one forward + backward on ``self.stream`` (graph replay when captured).
        ``wait_current``: order the step behind the caller's current stream first (inputs or the
        upstream gradient were written there)."""
        with torch.cuda.device(self.dev):
            if wait_current:
                self._stream.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(self._stream):
                if self._graph is not None:
                    self._graph.replay()
                else:
                    self._enqueue(self._stream, self._side)

    def capture_repeated(self, n: int) -> "torch.cuda.CUDAGraph":
        """ONE CUDA graph holding ``n`` whole steps back to back on this step's stream (the launch sequence
        of run(), n times).  A driver that issues many steps per host call is no longer paced by the host's
        graph-launch rate (one launch per ~50 us step per stream otherwise)."""
        with torch.cuda.device(self.dev):
            self._stream.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._stream):
                cur = torch.cuda.current_stream(self.dev)
                for _ in range(n):
                    self._enqueue(cur, self._side)
        return g

    def replay(self, graph) -> None:
        with torch.cuda.device(self.dev), torch.cuda.stream(self._stream):
            graph.replay()

    def run_cached_plan(self) -> None:
        """The same step with the plan of the LAST run() reused: feature staging + forward + backward only.
        This is evaluation with a fixed camera rig (patch.static_calibration): the index
        tables depend on the calibration alone."""
        with torch.cuda.device(self.dev), torch.cuda.stream(self._stream):
            if self._graph is None:
                self._enqueue(self._stream, self._side, with_plan=False)
                return
            if self._graph_cached is None:
                self._enqueue(self._stream, self._side, with_plan=False)      # warm
                self._stream.synchronize()
                self._graph_cached = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph_cached, stream=self._stream):
                    self._enqueue(torch.cuda.current_stream(self.dev), self._side, with_plan=False)
            self._graph_cached.replay()

    def load(self, tensors: Dict[str, torch.Tensor]) -> None:
        """Copy device or host tensors into the step's input buffers (on ``self.stream``)."""
        with torch.cuda.stream(self._stream):
            for k, v in tensors.items():
                self.inputs[k].copy_(v.reshape(self.in_shapes[k]), non_blocking=True)


class HostPipeline:
    """Feeds LiftSplatStep objects from pinned host memory, ``depth`` steps in flight.

    Every slot owns a pinned input block and a pinned output block and ONE CUDA graph holding the
    whole round trip: host->device copy of [calibration | depth | feat], the captured step, and
    the device->host copy of [d_depth | d_feat].  A data loader writes a batch straight into
    ``input_block(k)`` (laid out by ``pack``), ``submit()`` replays the slot's graph and
    ``collect()`` waits for the oldest step in flight and returns its host results.  With several
    slots the copies of one step overlap the kernels of the others (each slot has its own stream).
    ``submit(host_inputs)`` first packs / copies the given tensors into the slot's block (host
    memcpy; convenient, slower).
    """

    def __init__(self, make_step, depth: int = 2, graph_io: bool = True):
        self.slots = []
        for _ in range(depth):
            st: LiftSplatStep = make_step()
            h_in = torch.empty(st.in_block.numel(), dtype=torch.float32).pin_memory()
            h_out = torch.empty(st.out_block.numel(), dtype=torch.float32).pin_memory()
            slot = {"step": st, "h_in": h_in, "h_out": h_out, "done": torch.cuda.Event(), "busy": False,
                    "graph": None}
            self.slots.append(slot)
            self._round_trip(slot)                       # warm run outside capture
            st.stream.synchronize()
            if st._graph is not None and graph_io:       # the step itself is graph-captured: capture the I/O too
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=st.stream):
                    self._round_trip(slot, capturing=True)
                slot["graph"] = g
        self._next, self._oldest = 0, 0
        s0 = self.slots[0]["step"]
        self.h2d_bytes = s0.in_block.numel() * 4
        self.d2h_bytes = s0.out_block.numel() * 4

    @staticmethod
    def _round_trip(slot, capturing: bool = False) -> None:
        st: LiftSplatStep = slot["step"]
        stream = torch.cuda.current_stream(st.dev) if capturing else st.stream
        with torch.cuda.stream(stream):
            st.in_block.copy_(slot["h_in"], non_blocking=True)         # one H2D
            st._enqueue(stream, st._side)
            slot["h_out"].copy_(st.out_block, non_blocking=True)       # one D2H

    def input_block(self, k: int) -> torch.Tensor:
        """Pinned input block of slot k (layout: LiftSplatStep.in_block, see pack())."""
        return self.slots[k]["h_in"]

    def pack(self, host_inputs: Dict[str, torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Pack one step's inputs into a pinned block laid out like LiftSplatStep.in_block."""
        st: LiftSplatStep = self.slots[0]["step"]
        if out is None:
            out = torch.empty(st.in_block.numel(), dtype=torch.float32).pin_memory()
        for k, (o, n) in st._layout.items():
            out[o:o + n].copy_(host_inputs[k].reshape(-1))
        return out

    def submit(self, host_inputs=None) -> int:
        """Start one step on the next free slot and return the slot index.  ``host_inputs``: None
        (the slot's input block was filled in place), a dict of host tensors (rots, trans, intrins,
        post_rots, post_trans, depth, feat) or one block from pack()."""
        k = self._next
        slot = self.slots[k]
        if slot["busy"]:
            raise RuntimeError("pipeline full: collect() first")
        st: LiftSplatStep = slot["step"]
        if isinstance(host_inputs, torch.Tensor):
            if host_inputs.data_ptr() != slot["h_in"].data_ptr():
                slot["h_in"].copy_(host_inputs)
        elif host_inputs is not None:
            self.pack(host_inputs, slot["h_in"])
        if slot["graph"] is not None:
            with torch.cuda.stream(st.stream):
                slot["graph"].replay()
        elif st._graph is not None:
            with torch.cuda.stream(st.stream):
                st.in_block.copy_(slot["h_in"], non_blocking=True)
                st._graph.replay()
                slot["h_out"].copy_(st.out_block, non_blocking=True)
        else:
            self._round_trip(slot)
        slot["done"].record(st.stream)
        slot["busy"] = True
        self._next = (k + 1) % len(self.slots)
        return k

    def collect(self) -> Dict[str, torch.Tensor]:
        slot = self.slots[self._oldest]
        if not slot["busy"]:
            raise RuntimeError("nothing in flight")
        slot["done"].synchronize()
        slot["busy"] = False
        self._oldest = (self._oldest + 1) % len(self.slots)
        st: LiftSplatStep = slot["step"]
        n = st.ddepth.numel()
        return {"d_depth": slot["h_out"][:n].view(st.ddepth.shape),
                "d_feat": slot["h_out"][n:].view(st.dfeat.shape)}

    def in_flight(self) -> int:
        return sum(1 for s in self.slots if s["busy"])
