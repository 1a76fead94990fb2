"""Flow-export driver: the workload the SLIM hot path serves (reference ``liso/slim/experiment.py``).

* frame pairs are independent; the reference shards them over worker processes with
  ``sample_idx % world_size == worker_id`` (``experiment.py:330-332,351-353``) -> :func:`shard_indices`
* per pair it keeps the last-iteration BEV ``static_flow`` (H,W,2) and ``dynamicness`` (H,W) of both
  directions and writes them with ``np.savez_compressed`` (``experiment.py:391-399,459-471``)
  -> :func:`export_arrays`, :func:`save_npz`
* there is no communication in the reference; here one process drives one B200 and a single
  collective at the end sums counters / timings (NCCL over NVLink on GPUs, gloo in CPU tests)
  -> :func:`reduce_counters`
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Optional

import numpy as np
import torch


def shard_indices(n_samples: int, world_size: int, worker_id: int) -> List[int]:
    """Indices this worker owns: the reference's modulo rule (``experiment.py:330-332``)."""
    if world_size < 1 or not (0 <= worker_id < world_size):
        raise ValueError("bad shard spec world_size=%d worker_id=%d" % (world_size, worker_id))
    return [i for i in range(n_samples) if i % world_size == worker_id]


def export_arrays(preds_fw, preds_bw, threshold, batch_index: int = 0) -> Dict[str, torch.Tensor]:
    """The tensors the reference saves for one pair (``experiment.py:391-402``), still on device."""
    fw, bw = preds_fw[-1].modified_network_output, preds_bw[-1].modified_network_output
    return {
        "bev_raw_flow_t0_t1": fw.static_flow[batch_index],
        "bev_raw_flow_t1_t0": bw.static_flow[batch_index],
        "bev_dynamicness_t0_t1": fw.dynamicness[batch_index],
        "bev_dynamicness_t1_t0": bw.dynamicness[batch_index],
        "static_threshold": torch.as_tensor(threshold),
    }


def save_npz(target_file: str, arrays: Dict[str, torch.Tensor], bev_range_m, skip_existing: bool = False) -> bool:
    """Write one pair like ``slim_inference_and_save_result`` (``experiment.py:459-471``)."""
    if skip_existing and os.path.exists(target_file):
        return False
    out = {k: v.detach().cpu().numpy() for k, v in arrays.items()}
    out["bev_range_m"] = np.asarray(bev_range_m)
    os.makedirs(os.path.dirname(os.path.abspath(target_file)), exist_ok=True)
    np.savez_compressed(target_file, **out)
    return True


# directions of one exported sample: the t0 <-> t1 pair (Waymo / AV2), or the three pairs of the KITTI / nuScenes export
# (experiment.py:386-456); saved as bev_raw_flow_<dir> (H, W, 2) and bev_dynamicness_<dir> (H, W)
DIRECTIONS_PAIR = ("t0_t1", "t1_t0")
DIRECTIONS_TRIPLE = ("t0_t1", "t1_t0", "t0_t2", "t2_t0", "t1_t2", "t2_t1")


def export_keys(n_directions: int):
    dirs = {2: DIRECTIONS_PAIR, 6: DIRECTIONS_TRIPLE}[n_directions]
    return ["bev_raw_flow_" + d for d in dirs] + ["bev_dynamicness_" + d for d in dirs]


class AsyncNpzWriter:
    """the reference writes one ``np.savez_compressed`` file per pair on the thread that drives the
    GPU (``experiment.py:459-471``); at hundreds of pairs/s zlib would be the bottleneck, so the files are written
    by a pool of worker threads (zlib releases the GIL) while the GPU computes the next batches.  Same schema:
    ``static_threshold``, ``bev_raw_flow_{t0_t1,t1_t0}`` (H,W,2) f32, ``bev_dynamicness_{t0_t1,t1_t0}`` (H,W) f32,
    ``bev_range_m`` -- what ``torch_dataset_commons.py:619-675`` reads back.

    ``submit`` copies the arrays it is given (the caller's pinned buffers are reused) and returns immediately;
    at most ``max_pending`` files are in flight (back-pressure instead of unbounded memory)."""

    KEYS = ("bev_raw_flow_t0_t1", "bev_raw_flow_t1_t0", "bev_dynamicness_t0_t1", "bev_dynamicness_t1_t0")

    def __init__(self, target_dir: str, bev_range_m, workers: Optional[int] = None, max_pending: int = 64,
                 skip_existing: bool = False, compressed: bool = True, unlink_after_write: bool = False):
        from concurrent.futures import ThreadPoolExecutor

        self.target_dir = target_dir
        self.bev_range_m = np.asarray(bev_range_m)
        self.skip_existing = skip_existing
        self.compressed = compressed
        self.unlink_after_write = unlink_after_write  # benchmarks: bound the disk use (the write itself still happens)
        self.bytes_written = 0
        self.max_pending = max_pending
        self.pool = ThreadPoolExecutor(max_workers=workers or max(1, (os.cpu_count() or 2) - 1))
        self.pending: List = []
        self.written = 0
        os.makedirs(target_dir, exist_ok=True)

    def target_file(self, sample_id: str) -> str:
        return os.path.join(self.target_dir, sample_id + ".npz")  # sample ids may contain sub-folders (waymo)

    def _write(self, path: str, arrays: Dict[str, np.ndarray]) -> str:
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        tmp = path + ".tmp.npz"
        (np.savez_compressed if self.compressed else np.savez)(tmp, **arrays)
        os.replace(tmp, path)  # never leave a truncated file behind
        return self._written(path)

    def _written(self, path: str) -> str:
        self.bytes_written += os.path.getsize(path)  # (racy += across workers is fine for a statistic)
        if self.unlink_after_write:
            os.unlink(path)
        return path

    def submit(self, sample_id: str, *args) -> bool:
        """Arrays of ONE sample (numpy or CPU tensors): ``submit(id, flow_fw, flow_bw, dyn_fw, dyn_bw, threshold)`` for a
        pair, or ``submit(id, flows (6), dyns (6), threshold)`` in `DIRECTIONS_TRIPLE` order for the KITTI / nuScenes
        triple.  Returns False when the file exists and is skipped."""
        path = self.target_file(sample_id)
        if self.skip_existing and os.path.exists(path):
            return False
        *tensors, static_threshold = args
        if len(tensors) == 2 and isinstance(tensors[0], (list, tuple)):
            tensors = list(tensors[0]) + list(tensors[1])
        arrays = {"static_threshold": np.array(float(static_threshold), dtype=np.float32)}
        for k, v in zip(export_keys(len(tensors) // 2), tensors):
            arrays[k] = np.array(v.numpy() if torch.is_tensor(v) else v, dtype=np.float32, copy=True)
        arrays["bev_range_m"] = self.bev_range_m
        while len(self.pending) >= self.max_pending:
            self.pending.pop(0).result()
            self.written += 1
        self.pending.append(self.pool.submit(self._write, path, arrays))
        return True

    def submit_batch(self, sample_ids, host_tensors, static_threshold) -> int:
        """``host_tensors`` as handed to ``ExportPipeline``'s consume callback: batched flows (B,H,W,2) of all directions,
        then the dynamicness maps (B,H,W) of all directions (2 + 2 tensors for a pair, 6 + 6 for a triple) -- or, from a
        pipeline with ``compress=True``, the :class:`~liso_b200.slim.npz_stream.EncodedBatch` of the same maps."""
        n = 0
        if hasattr(host_tensors, "member"):
            for b, sid in enumerate(sample_ids):
                n += bool(self.submit_encoded(sid, host_tensors, b, static_threshold))
            return n
        for b, sid in enumerate(sample_ids):
            n += bool(self.submit(sid, *[t[b] for t in host_tensors], static_threshold))
        return n

    def _write_blob(self, path: str, members) -> str:
        from .npz_stream import build_npz

        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        tmp = path + ".tmp.npz"
        with open(tmp, "wb") as f:
            f.write(build_npz(members))
        os.replace(tmp, path)
        return self._written(path)

    def submit_encoded(self, sample_id: str, encoded, b: int, static_threshold) -> bool:
        """Sample ``b`` of a batch whose maps the GPU has already deflated (``DeflateEncoder``): the worker thread only
        frames the streams as zip members -- no zlib pass over the maps (``experiment.py:459-471`` spends its time there)."""
        path = self.target_file(sample_id)
        if self.skip_existing and os.path.exists(path):
            return False
        keys = export_keys(len(encoded.shapes) // 2)
        members = [("static_threshold", np.array(float(static_threshold), dtype=np.float32))]
        for v, k in enumerate(keys):
            shape, stream, r = encoded.member(v, b)  # (copies the bytes out of the pinned download buffer)
            members.append((k, shape, np.float32, stream, r))
        members.append(("bev_range_m", self.bev_range_m))
        while len(self.pending) >= self.max_pending:
            self.pending.pop(0).result()
            self.written += 1
        self.pending.append(self.pool.submit(self._write_blob, path, members))
        return True

    def close(self) -> int:
        for f in self.pending:
            f.result()
            self.written += 1
        self.pending = []
        self.pool.shutdown(wait=True)
        return self.written


def reduce_counters(local: Dict[str, float], device: Optional[torch.device] = None) -> Dict[str, float]:
    """Sum per-rank counters over the default process group (no-op without one).

    Keys ending in ``_max`` are reduced with MAX (timings), everything else with SUM."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(local)
    keys = sorted(local)
    dev = device if device is not None else torch.device("cpu")
    sums = torch.tensor([float(local[k]) for k in keys if not k.endswith("_max")], dtype=torch.float64, device=dev)
    maxs = torch.tensor([float(local[k]) for k in keys if k.endswith("_max")], dtype=torch.float64, device=dev)
    if sums.numel():
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    if maxs.numel():
        dist.all_reduce(maxs, op=dist.ReduceOp.MAX)
    out, i, j = {}, 0, 0
    for k in keys:
        if k.endswith("_max"):
            out[k] = float(maxs[j])
            j += 1
        else:
            out[k] = float(sums[i])
            i += 1
    return out


def iterate_batches(indices: Iterable[int], batch_size: int) -> Iterable[List[int]]:
    cur: List[int] = []
    for i in indices:
        cur.append(i)
        if len(cur) == batch_size:
            yield cur
            cur = []
    if cur:
        yield cur


def _is_lazy(x) -> bool:
    """A tensor-to-be: an object with ``shape``, ``dtype`` (torch) and ``write_into(uint8 numpy view)`` -- a sample source that
    can produce its data straight into the pinned upload buffer (decode / transform and copy in one pass)."""
    return hasattr(x, "write_into") and hasattr(x, "shape") and hasattr(x, "dtype")


def _walk(sample, fn):
    """The sample structure (dict / list / tensor / anything else) with ``fn`` applied to every tensor (or lazy tensor)."""
    if torch.is_tensor(sample) or _is_lazy(sample):
        return fn(sample)
    if isinstance(sample, dict):
        return {k: _walk(v, fn) for k, v in sample.items()}
    if isinstance(sample, (list, tuple)):
        return [_walk(v, fn) for v in sample]
    return sample


class PinnedArena:
    """A few grow-only pinned host buffers that whole batches are packed into by the loader thread: the upload of a batch
    is then ONE host-to-device copy (instead of one per tensor, each from a freshly pinned allocation), and no pinned memory
    is allocated in steady state.  A buffer goes back to the pool together with the event of the copy that reads it."""

    ALIGN = 256

    def __init__(self, buffers: int = 4):
        import queue

        self.free: "queue.Queue" = queue.Queue()
        for _ in range(buffers):
            self.free.put({"buf": None, "evt": None})

    def take(self, nbytes: int):
        slot = self.free.get()
        if slot["evt"] is not None:
            slot["evt"].synchronize()  # the copy that last read this buffer
            slot["evt"] = None
        if slot["buf"] is None or slot["buf"].numel() < nbytes:
            buf = torch.empty(max(nbytes + nbytes // 4, 1 << 20), dtype=torch.uint8)
            slot["buf"] = buf.pin_memory() if torch.cuda.is_available() else buf
        return slot

    def give_back(self, slot, evt=None):
        slot["evt"] = evt
        self.free.put(slot)

    def pack(self, frames, executor=None):
        """``frames``: tuple of (collated) sample dictionaries with host tensors -> :class:`PackedBatch`.  ``executor``: copy
        the tensors on its threads (the copies release the GIL; one thread moves ~6 GB/s into pinned memory)."""
        todo, off = [], 0

        def plan(t):
            nonlocal off
            off = (off + self.ALIGN - 1) // self.ALIGN * self.ALIGN
            n_el = int(np.prod(t.shape, dtype=np.int64))
            spec = (off, tuple(t.shape), t.dtype, n_el * torch.empty((), dtype=t.dtype).element_size())
            todo.append((spec, t))
            off += spec[3]
            return spec

        layout = tuple(_walk(f, plan) for f in frames)
        slot = self.take(off)
        raw = slot["buf"].numpy()  # (numpy copies: no intra-op thread team per copy, the GIL is released while bytes move)

        def copy(job):
            (o, shape, dtype, n), t = job
            if not n:
                return
            if _is_lazy(t):
                t.write_into(raw[o:o + n])
            else:
                np.copyto(raw[o:o + n], t.detach().contiguous().view(torch.uint8).reshape(-1).numpy())

        if executor is not None and len(todo) > 1:
            list(executor.map(copy, todo))
        else:
            for job in todo:
                copy(job)
        return PackedBatch(self, slot, off, layout)


class PackedBatch:
    """A batch packed into one pinned buffer (:meth:`PinnedArena.pack`); ``layout`` mirrors the sample dictionaries with
    every tensor replaced by (offset, shape, dtype, bytes)."""

    def __init__(self, arena, slot, nbytes, layout):
        self.arena, self.slot, self.nbytes, self.layout = arena, slot, nbytes, layout

    def __len__(self):
        return len(self.layout)

    def views(self, base: torch.Tensor):
        """The sample dictionaries as views of ``base`` (a uint8 buffer holding a copy of the packed bytes)."""
        def view(spec):
            o, shape, dtype, n = spec
            return base[o:o + n].view(dtype).view(shape)

        def walk(x):
            if isinstance(x, tuple) and len(x) == 4 and isinstance(x[2], torch.dtype):
                return view(x)
            if isinstance(x, dict):
                return {k: walk(v) for k, v in x.items()}
            if isinstance(x, list):
                return [walk(v) for v in x]
            return x

        return tuple(walk(f) for f in self.layout)


class ExportPipeline:
    """Double-buffered export loop for one GPU: the host->device upload of batch i+1 and the device->host download
    of batch i-1 run on a copy stream while batch i is computed on the caller's stream.  (The reference uploads,
    computes and ``.cpu().numpy()``s each sample strictly in sequence, ``experiment.py:363-402``.)

    ``batches``: iterable of ``(sample_data_t0, sample_data_t1)`` -- or ``(t0, t1, t2)`` for the KITTI / nuScenes triple,
    run as ONE ``SLIM.forward_triple`` pass in which every frame is encoded once -- with tensors in *pinned* host memory.
    ``consume(index, arrays)``: called with the exported tensors of a batch as pinned host tensors once they have
    arrived (they are reused ``depth`` batches later: copy or write them out inside the callback).
    """

    KEYS = ("pcl_full_no_ground_ta", "pcl_ta")

    def __init__(self, model, device, amp_ctx=None, depth: int = 3, compress: bool = False):
        """``depth``: batches the host may run ahead of the results it has consumed (number of upload / download
        slots); >= 2.  A deeper pipeline absorbs host-side hiccups (the launch thread is the critical resource).
        ``compress``: the exported maps are DEFLATE-compressed on the GPU right behind the forward pass (SURVEY 8f.3,
        ``npz_stream.DeflateEncoder``) and only the compressed streams are downloaded; ``consume`` then receives an
        ``EncodedBatch`` (views in the order flows of all directions, dynamicness of all directions)."""
        self.model, self.device = model, device
        self.compress = bool(compress)
        self.encoder = None
        self.downloaded_bytes = 0
        # where the launch thread's time goes (seconds, summed over `run` calls): waiting for the next batch from the
        # loader, enqueueing a batch (Python + launches), blocked on the device (member table / downloads), in `consume`
        self.host_s = {"input_wait": 0.0, "launch": 0.0, "gpu_wait": 0.0, "consume": 0.0, "batches": 0}
        if hasattr(model, "outputs_alias_static_buffers"):
            model.outputs_alias_static_buffers = True  # this loop takes its own packed copies of what it exports (see run)
        self.copy_stream = torch.cuda.Stream(device=device)
        self.enc_stream = torch.cuda.Stream(device=device)  # the DEFLATE encoder of batch i runs under the front end of batch i + 1
        self.amp_ctx = amp_ctx
        self.depth = max(2, int(depth))
        self._out_bufs = [None] * self.depth
        self._up_bufs = {}

    def _stage(self, slot: int, name: str, host: torch.Tensor) -> torch.Tensor:
        """Copy ``host`` into a grow-only device buffer owned by (slot, name): no allocator traffic, no cross-stream
        block reuse in steady state."""
        key = (slot, name)
        buf = self._up_bufs.get(key)
        n = host.numel()
        if buf is None or buf.numel() < n or buf.dtype != host.dtype:
            buf = torch.empty(max(n, 1), dtype=host.dtype, device=self.device)
            self._up_bufs[key] = buf
        dst = buf[:n].view(host.shape)
        dst.copy_(host, non_blocking=True)
        return dst

    def _upload(self, sample, slot: int, tag: str):
        out = {}
        for k, v in sample.items():
            if isinstance(v, dict):
                out[k] = {kk: self._stage(slot, "%s/%s/%s" % (tag, k, kk), vv) for kk, vv in v.items()}
            elif isinstance(v, (list, tuple)):
                out[k] = [self._stage(slot, "%s/%s/%d" % (tag, k, i), t) for i, t in enumerate(v)]
            elif torch.is_tensor(v):
                out[k] = self._stage(slot, "%s/%s" % (tag, k), v)
            else:
                out[k] = v
        return out

    def _upload_batch(self, batch, slot: int):
        """Stage one batch on the device (called on the copy stream): a :class:`PackedBatch` is ONE copy into the slot's
        grow-only device buffer, the sample dictionaries are views of it; a tuple of sample dictionaries is copied tensor
        by tensor."""
        if isinstance(batch, PackedBatch):
            key = (slot, "packed")
            buf = self._up_bufs.get(key)
            if buf is None or buf.numel() < batch.nbytes:
                buf = torch.empty(max(batch.nbytes + batch.nbytes // 4, 1 << 20), dtype=torch.uint8, device=self.device)
                self._up_bufs[key] = buf
            buf[:batch.nbytes].copy_(batch.slot["buf"][:batch.nbytes], non_blocking=True)
            evt = torch.cuda.Event()
            evt.record(torch.cuda.current_stream(self.device))
            batch.arena.give_back(batch.slot, evt)  # the pinned buffer is free again once this copy has run
            return batch.views(buf)
        return tuple(self._upload(s, slot, "t%d" % t) for t, s in enumerate(batch))

    def run(self, batches, consume=None):
        import contextlib
        import time

        clock, hs = time.perf_counter, self.host_s
        cur = torch.cuda.current_stream(self.device)
        it = iter(batches)
        downloads = []  # (index, host tensors, event)
        D = self.depth
        done_evt = [None] * D  # compute-finished event of the batch that last used each upload slot
        nxt = next(it, None)
        staged = None
        if nxt is not None:
            self.copy_stream.wait_stream(cur)
            with torch.cuda.stream(self.copy_stream):
                staged = self._upload_batch(nxt, 0)
                up_evt = torch.cuda.Event()
                up_evt.record(self.copy_stream)
        idx = 0
        n_done = 0
        while staged is not None:
            t_launch = clock()
            cur.wait_event(up_evt)
            ctx = self.amp_ctx() if self.amp_ctx else contextlib.nullcontext()
            with torch.no_grad(), ctx:
                if any("pcl_ta" not in s for s in staged):  # raw scans: decoder inputs prepared on the device (8f.4)
                    from ..datasets import preprocess_scans

                    staged = tuple(s if "pcl_ta" in s else preprocess_scans(s["pcl_full_w_ground_ta"], self.model.cfg) for s in staged)
                if len(staged) == 3:
                    by_dir = self.model.forward_triple(*staged)
                    mods = [by_dir[d][-1].modified_network_output for d in DIRECTIONS_TRIPLE]
                else:
                    pf, pb = self.model(staged[0], staged[1], None)
                    mods = [pf[-1].modified_network_output, pb[-1].modified_network_output]
            if self.compress:
                # the encoder reads the (strided) maps where the decoder left them, in stream order before the next forward
                # overwrites the graph's static outputs; the member table follows on the same stream (a few hundred bytes)
                if self.encoder is None:
                    from .npz_stream import DeflateEncoder

                    self.encoder = DeflateEncoder(self.device, slots=D)
                views = [m.static_flow for m in mods] + [m.dynamicness for m in mods]
                net = getattr(self.model, "raft_network", None)
                if net is not None and getattr(net, "last_forward_was_graph", False) and getattr(self.model, "outputs_alias_static_buffers", False):
                    # the maps are the graph's static outputs: encode on a side stream, under the pre-processing and the
                    # pillar encoder of the next batch; only the next graph replay (which overwrites them) waits for it
                    fwd_done = torch.cuda.Event()
                    fwd_done.record(cur)
                    self.enc_stream.wait_event(fwd_done)
                    with torch.cuda.stream(self.enc_stream):
                        self.encoder.encode(views, slot=idx % D)
                        self.encoder.start_download(idx % D)
                        enc_done = torch.cuda.Event()
                        enc_done.record(self.enc_stream)
                    net.pre_replay_event = enc_done
                else:
                    self.encoder.encode(views, slot=idx % D)
                    self.encoder.start_download(idx % D)
                outs = None
            else:
                # packed copies: the predictions may be views of the CUDA graph's static outputs, which the next forward
                # overwrites while this batch is still being downloaded
                outs = [m.static_flow.contiguous() for m in mods] + [m.dynamicness.contiguous() for m in mods]
            done = torch.cuda.Event()
            done.record(cur)
            done_evt[idx % D] = done
            t_in = clock()
            hs["launch"] += t_in - t_launch
            hs["batches"] += 1
            # stage the next batch and download this one on the copy stream, behind the compute of this batch
            nxt = next(it, None)
            hs["input_wait"] += clock() - t_in
            with torch.cuda.stream(self.copy_stream):
                if nxt is not None:
                    up_slot = (idx + 1) % D
                    if done_evt[up_slot] is not None:  # the batch that last read this slot's buffers must be through
                        self.copy_stream.wait_event(done_evt[up_slot])
                    staged = self._upload_batch(nxt, up_slot)
                    up_evt = torch.cuda.Event()
                    up_evt.record(self.copy_stream)
                else:
                    staged = None
                slot = idx % D
                if not self.compress:
                    self.copy_stream.wait_event(done)
                    if self._out_bufs[slot] is None or any(b.shape != o.shape for b, o in zip(self._out_bufs[slot], outs)):
                        self._out_bufs[slot] = [torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in outs]
                    for dst, src in zip(self._out_bufs[slot], outs):
                        src.record_stream(self.copy_stream)
                        dst.copy_(src, non_blocking=True)
                    dl_evt = torch.cuda.Event()
                    dl_evt.record(self.copy_stream)
            if self.compress:
                # the previous batch has left the GPU by now (or is about to): its sizes are known -> download exactly its
                # compressed bytes on the copy stream, under the compute of the batch just enqueued
                if downloads and not downloads[-1][2]["begun"]:
                    self._begin(downloads[-1])
                downloads.append((idx, slot, {"begun": False}))
            else:
                downloads.append((idx, self._out_bufs[slot], dl_evt))
            # hand over the oldest batch once `depth - 1` newer ones are in flight (its pinned buffers come up for reuse)
            while len(downloads) > D - 1:
                n_done += self._hand_over(downloads.pop(0), consume)
            idx += 1
        for d in downloads:
            n_done += self._hand_over(d, consume)
        cur.wait_stream(self.enc_stream)  # whoever uses the model next (on `cur`) must not overwrite maps still being encoded
        net = getattr(self.model, "raft_network", None)
        if net is not None and hasattr(net, "pre_replay_event"):
            net.pre_replay_event = None
        return n_done

    def _begin(self, d):
        import time

        t = time.perf_counter()
        self.downloaded_bytes += self.encoder.fetch_begin(d[1], self.copy_stream)
        d[2]["begun"] = True
        self.host_s["gpu_wait"] += time.perf_counter() - t

    def _hand_over(self, d, consume) -> int:
        import time

        t = time.perf_counter()
        if self.compress:
            if not d[2]["begun"]:
                self._begin(d)
                t = time.perf_counter()
            host = self.encoder.fetch_end(d[1])
        else:
            d[2].synchronize()
            host = d[1]
        t1 = time.perf_counter()
        self.host_s["gpu_wait"] += t1 - t
        if consume is not None:
            consume(d[0], host)
        self.host_s["consume"] += time.perf_counter() - t1
        return 1


def collate_pairs(pairs):
    """Batch un-batched pairs the way the reference's collate function does (``torch_dataset_commons.py:380-401``): the
    network clouds stay a list; ``pcl_ta`` is padded to the longest cloud with NaN points, ``-1`` pillar coordinates
    and a validity mask.  Samples that are raw scans only (``{"pcl_full_w_ground_ta", "raw_scan": True}``) are passed on
    as a list of scans.  ``pairs``: list of ``(sample_t0, sample_t1)`` with ``pcl_full_no_ground_ta`` (N, C) and
    ``pcl_ta = {"pcl" (N', C), "pillar_coors" (N', 2)}`` per sample (host tensors)."""
    from torch.nn.utils.rnn import pad_sequence

    out = []
    for t in range(len(pairs[0])):
        samples = [p[t] for p in pairs]
        if "pcl_ta" not in samples[0]:  # raw scans: the device prepares the decoder inputs (ExportPipeline)
            out.append({"pcl_full_w_ground_ta": [s["pcl_full_w_ground_ta"] for s in samples], "raw_scan": True})
            continue
        pcl = pad_sequence([s["pcl_ta"]["pcl"] for s in samples], batch_first=True, padding_value=float("nan"))
        coors = pad_sequence([s["pcl_ta"]["pillar_coors"] for s in samples], batch_first=True, padding_value=-1)
        batch = {"pcl_full_no_ground_ta": [s["pcl_full_no_ground_ta"] for s in samples],
                 "pcl_ta": {"pcl": pcl, "pillar_coors": coors, "pcl_is_valid": torch.logical_not(torch.isnan(pcl).sum(-1) > 0)}}
        if all("raw_scan" in s for s in samples):
            batch["raw_scan"] = samples[0]["raw_scan"]
        out.append(batch)
    return tuple(out)


def _prefetched(it, depth: int):
    """Run the iterator ``it`` on a background thread, ``depth`` items ahead (the dataset access, collation and pinning of
    the next batches then overlap the launch thread, like the workers of the reference's DataLoader)."""
    import queue
    import threading

    q: "queue.Queue" = queue.Queue(maxsize=depth)
    end = object()

    def work():
        try:
            for item in it:
                q.put(item)
            q.put(end)
        except BaseException as e:  # hand the error to the consumer instead of dying silently
            q.put(e)

    threading.Thread(target=work, daemon=True).start()
    while True:
        item = q.get()
        if item is end:
            return
        if isinstance(item, BaseException):
            raise item
        yield item


def _pin(sample):
    def pin(t):
        if _is_lazy(t):
            t = t.materialize()
        return t.pin_memory() if torch.cuda.is_available() and not t.is_pinned() else t

    return {k: ([pin(t) for t in v] if isinstance(v, (list, tuple)) else {kk: pin(vv) for kk, vv in v.items()}
                if isinstance(v, dict) else (pin(v) if (torch.is_tensor(v) or _is_lazy(v)) else v)) for k, v in sample.items()}


def run_flow_export(model, dataset, target_dir: str, bev_range_m, *, world_size: int = 1, worker_id: int = 0,
                    batch_size: int = 8, device=None, skip_existing: bool = False, writer_workers: Optional[int] = None,
                    pipeline_factory=None, compress_on_gpu: Optional[bool] = None, loader_workers: int = 0,
                    unlink_after_write: bool = False, pad_last_batch: bool = True, pack_uploads: bool = True) -> Dict[str, float]:
    """The flow export of ``liso/slim/experiment.py:225-361,363-471`` for the t0 -> t1 pairs of one worker: this rank's
    share of the pairs (modulo rule), batched, through the double-buffered :class:`ExportPipeline`, written by
    :class:`AsyncNpzWriter` in the reference's ``.npz`` schema under ``target_dir/<sample_id>.npz``; one collective at
    the end sums the counters over the ranks.

    ``dataset``: ``len()`` and ``dataset[i] -> (sample_id, sample_t0, sample_t1)`` with un-batched host tensors (see
    :func:`collate_pairs`), or ``(sample_id, t0, t1, t2)`` for the KITTI / nuScenes export: then t0 -> t1, t0 -> t2 and
    t1 -> t2 are computed in one pass per batch (every frame encoded once) and the file holds all 12 maps
    (``experiment.py:404-456``).  ``compress_on_gpu`` (default: on for a CUDA device): the maps are deflated on the device and
    the writer threads only frame zip members (same files for ``np.load``); off: ``np.savez_compressed`` on host threads.  ``loader_workers``: > 0 moves dataset access, collation and
    pinning to a background thread (> 1: with that many threads fetching samples), like DataLoader workers.
    ``pack_uploads``: every batch is packed into one recycled pinned buffer and uploaded with one copy.
    ``pad_last_batch``: a ragged last batch is filled up with copies of its last sample (outputs discarded) so that no
    tensor shape changes.  Returns ``{"pairs", "files", "skipped", "elapsed_s_max"}`` over all ranks ("pairs"
    counts samples)."""
    import time

    if compress_on_gpu is None:  # default: the device-side writer whenever the stock pipeline drives a CUDA device
        compress_on_gpu = (pipeline_factory is None and torch.cuda.is_available() and device is not None
                           and torch.device(device).type == "cuda")
    writer = AsyncNpzWriter(target_dir, bev_range_m, workers=writer_workers, skip_existing=skip_existing,
                            unlink_after_write=unlink_after_write)
    mine = shard_indices(len(dataset), world_size, worker_id)
    ids_of_batch: List[List[str]] = []
    skipped = 0

    # batches packed into recycled pinned buffers (one upload per batch) when the stock pipeline drives a CUDA device
    arena = PinnedArena(4) if (pipeline_factory is None and pack_uploads and torch.cuda.is_available()) else None
    fetch_pool = None
    if loader_workers > 1:
        from concurrent.futures import ThreadPoolExecutor

        fetch_pool = ThreadPoolExecutor(loader_workers)

    loader_s = [0.0]

    def batches():
        nonlocal skipped
        for chunk in iterate_batches(mine, batch_size):
            t_load = time.perf_counter()
            items = list(fetch_pool.map(dataset.__getitem__, chunk)) if fetch_pool else [dataset[i] for i in chunk]
            if skip_existing:  # (experiment.py:380-382: an existing target file skips the forward as well)
                keep = [it for it in items if not os.path.exists(writer.target_file(it[0]))]
                skipped += len(items) - len(keep)
                items = keep
            if not items:
                continue
            ids_of_batch.append([it[0] for it in items])
            if pad_last_batch and len(items) < batch_size:
                # a ragged last batch would change every tensor shape (new cuDNN plans, a new CUDA graph for one batch):
                # fill it up with copies of its last sample; their outputs are never read (ids_of_batch has the real ones)
                items = items + [items[-1]] * (batch_size - len(items))
            frames = collate_pairs([tuple(it[1:]) for it in items])
            out = arena.pack(frames, fetch_pool) if arena is not None else tuple(_pin(d) for d in frames)
            loader_s[0] += time.perf_counter() - t_load
            yield out

    thr = float(model.moving_dynamicness_threshold.value())
    counts = {"pairs": 0, "files": 0}

    def consume(j, host):
        counts["pairs"] += len(ids_of_batch[j])
        counts["files"] += writer.submit_batch(ids_of_batch[j], host, thr)

    pipeline = (pipeline_factory(model, device) if pipeline_factory else ExportPipeline(model, device, compress=compress_on_gpu))
    t0 = time.perf_counter()
    pipeline.run(_prefetched(batches(), 3) if loader_workers > 0 else batches(), consume)
    writer.close()
    if fetch_pool:
        fetch_pool.shutdown()
    local = {"pairs": float(counts["pairs"]), "files": float(counts["files"]), "skipped": float(skipped),
             "file_bytes": float(writer.bytes_written), "d2h_bytes": float(getattr(pipeline, "downloaded_bytes", 0)),
             "elapsed_s_max": time.perf_counter() - t0}
    hs = getattr(pipeline, "host_s", None)
    if hs and hs.get("batches"):  # host-side time per batch (ms), summed over the ranks by reduce_counters: divide by the world size
        nb = float(hs["batches"])
        local.update({"host_ms_launch": 1e3 * hs["launch"] / nb, "host_ms_input_wait": 1e3 * hs["input_wait"] / nb,
                      "host_ms_gpu_wait": 1e3 * hs["gpu_wait"] / nb, "host_ms_consume": 1e3 * hs["consume"] / nb,
                      "host_ms_loader": 1e3 * loader_s[0] / nb})
    return reduce_counters(local, device if device is not None and torch.device(device).type == "cuda" else None)
