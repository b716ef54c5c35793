# -*- coding: utf-8 -*-
"""
The drop-in entry points: ``raytrace(config)``, ``raytrace_single(config)``,
``raytrace_mp(config)``, ``combine_raytrace(list)`` with the reference's
config dict in and its ``{'config','total','found','lost'}`` dict out
(reference ``xicsrt/xicsrt_raytrace.py:28-393``).

What differs from the reference is only *where* an iteration runs: the body of
``_raytrace_iter`` + ``_sort_raytrace`` (``xicsrt_raytrace.py:178-278``) is one
fused kernel launch (generate -> optic train -> bin) plus, when history is
kept, a replay of the selected found / lost ray ids that writes their
per-element states as struct-of-arrays.  Random numbers are Philox4x32-10
keyed by ``(random_seed, iteration)`` and counted by global ray id, so results
do not depend on how rays are partitioned over launches or GPUs (and are,
by design, a different stream than the reference's MT19937).

PyTorch only owns the device buffers; there is no CPU path.
"""
import atexit
import collections
import copy
import ctypes as C
import logging
import os

import numpy as np

from . import _lib as L
from . import config as xconfig
from . import scene as xscene

log = logging.getLogger('xicsrt_b200')

RAY_KEYS = ('origin', 'direction', 'mask', 'wavelength')
U64_MAX = (1 << 64) - 1


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError('xicsrt_b200 needs a CUDA device (B200); there is no CPU path.')
    return torch


def _skeleton(config):
    return {'config': config,
            'total': {'meta': {}, 'image': {}},
            'found': {'meta': {}, 'history': {}},
            'lost': {'meta': {}, 'history': {}}}


def shard_range(n, rank, world):
    """Contiguous global ray-id range of one rank: [begin, begin + count)."""
    begin = (n * rank) // world
    end = (n * (rank + 1)) // world
    return begin, end - begin


def allreduce_packed(packed):
    """
    The one data-path collective of a multi-GPU iteration: sum of the packed int64 buffer
    ``[counts | image_0 | image_1 | ...]`` over ranks (NCCL over NVLink on the GPUs; gloo in
    the CPU tests).  Integer sums, so the result is independent of the reduction order.
    """
    import torch.distributed as dist
    dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    return packed


class HostRandom:
    """
    Host-side draws (Poisson ray counts, plasma bundle centres): numpy Philox keyed by the run
    seed and a stream number (0 = run set-up, 1 + i = iteration i), so every rank of a multi-GPU
    run builds the same tables.
    """

    def __init__(self, seed, stream=0):
        self.gen = np.random.Generator(np.random.Philox(key=[int(seed) & U64_MAX, int(stream) & U64_MAX]))

    def poisson(self, lam):
        return int(self.gen.poisson(lam))

    def poisson_array(self, lam):
        return self.gen.poisson(lam).astype(np.int64)

    def uniform(self, lo, hi, n):
        return self.gen.uniform(lo, hi, n)


class Tracer:
    """
    One prepared run on one GPU: elements prepared on the host, scene uploaded,
    device buffers allocated once and reused over iterations.
    """

    def __init__(self, config, seed, rank=0, world=1, device=None):
        torch = _torch()
        self.torch = torch
        self.rank, self.world = rank, world
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.seed = int(seed) & U64_MAX
        self.host_rng = HostRandom(self.seed)

        (self.config, self.source_name, self.source_param, self.source_filters,
         self.optics) = xscene.prepare(config, poisson=self.host_rng.poisson)
        self._sections = None      # set by the scene cache: pristine copies of the defaulted element configs
        self.is_plasma = self.source_param['_kind'].startswith('plasma')
        self.scene = None
        self._upload(0)
        self.lib = self.scene.lib
        self.n_elem = 1 + len(self.layout.optic_names)

        i64 = torch.int64
        # one packed buffer [counts | images] so that a multi-GPU run reduces it with one call
        self.packed = torch.zeros(self.n_elem + max(self.layout.n_pixels, 1), dtype=i64, device=self.device)
        self.scalars = torch.zeros(2, dtype=i64, device=self.device)     # found_count, lost_count
        self.found_ids = None
        self.lost_ids = None
        self.lost_keys = None

    # ------------------------------------------------------------------
    def _upload(self, iteration):
        """Flatten and upload the scene (once per run); plasma sources then get their first bundle table."""
        desc, self.layout, keep = xscene.flatten(self.source_name, self.source_param, self.source_filters,
                                                 self.optics)
        with self.torch.cuda.device(self.device):
            self.scene = xscene.DeviceScene(desc, self.layout)
        self.upload_bytes = keep.nbytes()      # host -> device bytes of the tables behind the descriptor
        del keep
        self.n_rays = self.layout.n_rays
        self.bundles = None
        if self.is_plasma:
            from . import plasma
            self.bundles = plasma.DeviceBundles(self.torch, self.device, self.source_param, self.source_filters,
                                                self.scene.lib)
            self._new_bundles(iteration)

    def _new_bundles(self, iteration):
        """Bundle centres, plasma parameters and ray counts of one iteration, built on the device."""
        self.n_rays = self.bundles.generate(self.seed, (1 << 32) + iteration)
        self.layout.n_rays = self.n_rays
        self.scene.set_bundles(self.bundles.table, self.bundles.end, self.n_rays)
        if self.bundles.voigt_x is not None:
            self.scene.set_bundle_tables(self.bundles.voigt_x, self.bundles.voigt_cdf)

    def begin_iteration(self, iteration):
        """
        The reference re-runs setup_bundles in every generate_rays call
        (_XicsrtPlasmaGeneric.py:384-393): new bundle centres and ray counts each iteration.
        """
        if self.is_plasma and iteration > 0:
            self._new_bundles(iteration)

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _outputs(self, keep_images, found_cap=0, lost_cap=0, lost_threshold=0, want_lists=False, packed=None, bitmaps=None):
        packed = self.packed if packed is None else packed
        out = L.XrtOutputs()
        out.counts = packed.data_ptr()
        out.images = packed.data_ptr() + 8 * self.n_elem if (keep_images and self.layout.n_pixels) else None
        if bitmaps is not None:
            out.found_bits, out.lost_bits, out.bits_begin = bitmaps
            out.lost_threshold = lost_threshold
        if want_lists:
            torch = self.torch
            if self.found_ids is None or self.found_ids.numel() < found_cap:
                self.found_ids = torch.empty(max(found_cap, 1), dtype=torch.int64, device=self.device)
            if self.lost_ids is None or self.lost_ids.numel() < lost_cap:
                self.lost_ids = torch.empty(max(lost_cap, 1), dtype=torch.int64, device=self.device)
                self.lost_keys = torch.empty(max(lost_cap, 1), dtype=torch.int64, device=self.device)
            out.found_ids = self.found_ids.data_ptr()
            out.found_count = self.scalars.data_ptr()
            out.found_capacity = found_cap
            out.lost_ids = self.lost_ids.data_ptr()
            out.lost_keys = self.lost_keys.data_ptr()
            out.lost_count = self.scalars.data_ptr() + 8
            out.lost_capacity = lost_cap
            out.lost_threshold = lost_threshold
        return out

    def trace(self, stream_id, keep_images=True, ray_begin=None, ray_count=None, zero=True,
              found_cap=0, lost_cap=0, lost_threshold=0, want_lists=False, packed=None, bitmaps=None):
        """
        Enqueue one fused generate->trace->bin launch for this rank's ray range (asynchronous).  ``packed`` = another
        [counts | images] buffer of the same shape to accumulate into (double buffering against an all-reduce in flight).
        """
        if ray_begin is None:
            ray_begin, ray_count = shard_range(self.n_rays, self.rank, self.world)
        if zero:
            (self.packed if packed is None else packed).zero_()
            self.scalars.zero_()
        out = self._outputs(keep_images, found_cap, lost_cap, lost_threshold, want_lists, packed, bitmaps)
        with self.torch.cuda.device(self.device):
            L.check(self.lib.xrt_trace(self.scene.handle, self.seed, int(stream_id), int(ray_begin), int(ray_count),
                                       C.byref(out), self._stream()))
        return ray_begin, ray_count

    def history(self, stream_id, ids, out=None, rows=False):
        """
        Replay the given global ray ids (int64 device tensor); returns (rays[E,7,n], mask[E,n]) on
        device.  ``out`` = (rays, mask) tensors of a previous call may be passed to reuse their memory.
        ``rows=True``: the 7 n doubles of an element hold the reference's row arrays instead of planes --
        origin (n,3), direction (n,3), wavelength (n,) back to back (XRT_HIST_ROWS; capacity is exactly n).
        """
        torch = self.torch
        n = int(ids.numel())
        cap = max(n, 1)
        if out is not None and out[0].shape[2] >= cap and out[0].shape[0] == self.n_elem:
            rays, mask = out
            cap = rays.shape[2]
        else:
            rays = torch.empty((self.n_elem, 7, cap), dtype=torch.float64, device=self.device)
            mask = torch.empty((self.n_elem, cap), dtype=torch.uint8, device=self.device)
        if n:
            h = L.XrtHistory()
            h.rays, h.mask, h.capacity = rays.data_ptr(), mask.data_ptr(), cap
            h.layout = L.HIST_ROWS if rows else L.HIST_PLANES
            with torch.cuda.device(self.device):
                L.check(self.lib.xrt_trace_history(self.scene.handle, self.seed, int(stream_id), ids.data_ptr(),
                                                   0, n, C.byref(h), self._stream()))
        return rays[:, :, :n], mask[:, :n]

    # ------------------------------------------------------------------
    def counts_and_images(self, keep_images=True):
        """Host copies of the packed counters: ({name: num_out}, {name: image or None})."""
        host = self.packed.cpu().numpy()
        names = self.layout.element_names
        meta = {name: int(host[i]) for i, name in enumerate(names)}
        image = {}
        if keep_images:
            for name in self.layout.optic_names:
                spec = self.layout.images[name]
                if spec is None:
                    image[name] = None
                else:
                    off, nx, ny = spec
                    image[name] = host[self.n_elem + off:self.n_elem + off + nx * ny].astype(np.float64).reshape(nx, ny)
        return meta, image

    def allreduce(self, packed=None):
        """Sum counters + images over ranks (one collective on the packed buffer), on the current stream."""
        if self.world > 1:
            allreduce_packed(self.packed if packed is None else packed)

    def select_ids(self, stream_id, max_lost, keep_images=True):
        """
        One launch that also marks the found rays and the candidates of the lost sample in two bitmaps indexed by
        ray id; the library's count / scan / emit kernels turn them into id lists (``xrt_bits_to_ids``) and
        ``xrt_lost_select`` picks the ``max_lost`` candidates with the smallest Philox keys.  Returns
        (found_ids, lost_ids) as int64 device tensors, both in ascending id order: the found rays are in the
        reference's ray order, the lost sample is a uniform random subset as ``_sort_raytrace`` draws with a shuffle
        (xicsrt_raytrace.py:262-266).  A second launch happens only when the candidate sample turns out too small
        (almost every ray found).
        """
        torch = self.torch
        begin, count = shard_range(self.n_rays, self.rank, self.world)
        n_words = (count + 31) // 32
        if getattr(self, 'bits', None) is None or self.bits.numel() < 2 * n_words:
            self.bits = torch.empty(max(2 * n_words, 2), dtype=torch.int32, device=self.device)
        bits = self.bits[:2 * n_words]
        cnt = torch.zeros(1, dtype=torch.int64, device=self.device)
        want = 2 * max_lost + 10 * int(np.sqrt(max_lost)) + 64
        prob = min(1.0, want / max(count, 1))

        def ids_of(ptr, capacity):
            buf = torch.empty(max(capacity, 1), dtype=torch.int64, device=self.device)
            with torch.cuda.device(self.device):
                L.check(self.lib.xrt_bits_to_ids(ptr, count, begin, buf.data_ptr(), capacity, cnt.data_ptr(), self._stream()))
            return buf, int(cnt.cpu()[0])

        while True:
            thr = U64_MAX if prob >= 1.0 else int(prob * float(1 << 64))
            bits.zero_()
            self.trace(stream_id, keep_images, begin, count, True, lost_threshold=thr,
                       bitmaps=(bits.data_ptr(), bits.data_ptr() + 4 * n_words, begin))
            n_found = int(self.packed[self.n_elem - 1].cpu())
            n_lost = count - n_found
            cap = min(n_lost, int(prob * count * 1.5) + 4096)
            cand, n_cand = ids_of(bits.data_ptr() + 4 * n_words, cap)
            if n_cand > cap:
                cand, n_cand = ids_of(bits.data_ptr() + 4 * n_words, n_cand)
            if n_cand < min(max_lost, n_lost) and prob < 1.0:
                prob = min(1.0, 4.0 * prob * max(1.0, min(max_lost, n_lost) / max(n_cand, 1)))
                continue
            break
        found, n_f = ids_of(bits.data_ptr(), n_found)
        assert n_f == n_found, (n_f, n_found)
        m = min(max_lost, n_cand)
        lost = torch.empty(max(m, 1), dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            L.check(self.lib.xrt_lost_select(self.seed, int(stream_id), cand.data_ptr(), n_cand, m, lost.data_ptr(),
                                             cnt.data_ptr(), self._stream()))
        return found[:n_found], lost[:m]

    def rebind(self, config, seed):
        """
        A cached tracer (same elements, see :func:`_scene_key`) takes over another call: new seed, and the call's own
        config object completed with copies of the defaulted element configs (what ``prepare`` writes back).
        """
        self.seed = int(seed) & U64_MAX
        self.host_rng = HostRandom(self.seed)
        for k, v in self._sections.items():
            config[k] = copy.deepcopy(v)
        self.config = config

    def close(self):
        self.scene.close()


def _history_dicts(names, rays, mask):
    """Device SoA [E,7,n] / [E,n] -> {element: {'origin','direction','wavelength','mask'}} on the host."""
    rays = rays.cpu().numpy()
    mask = mask.cpu().numpy()
    out = {}
    for e, name in enumerate(names):
        out[name] = {
            'origin': np.ascontiguousarray(rays[e, 0:3, :].T),
            'direction': np.ascontiguousarray(rays[e, 3:6, :].T),
            'mask': mask[e].astype(np.bool_),
            'wavelength': np.ascontiguousarray(rays[e, 6, :]),
        }
    return out


def rows_views(rays, mask, e, lo, hi):
    """
    Views of rows [lo, hi) of element e in a rows-layout history (XRT_HIST_ROWS): rays [E,7,n] holds, per element,
    origin (n,3), direction (n,3), wavelength (n,) back to back.  Works on torch tensors and numpy arrays alike.
    """
    n = rays.shape[2]
    flat = rays[e].reshape(-1)
    return {'origin': flat[0:3 * n].reshape(n, 3)[lo:hi],
            'direction': flat[3 * n:6 * n].reshape(n, 3)[lo:hi],
            'wavelength': flat[6 * n:7 * n][lo:hi],
            'mask': mask[e][lo:hi]}


def to_host(torch, rays, mask):
    """Device -> host through pinned memory (torch's caching host allocator), one synchronisation; numpy views."""
    if rays.device.type != 'cuda':
        return rays.numpy(), mask.numpy()
    h_rays = torch.empty(rays.shape, dtype=rays.dtype, pin_memory=True)
    h_mask = torch.empty(mask.shape, dtype=mask.dtype, pin_memory=True)
    h_rays.copy_(rays, non_blocking=True)
    h_mask.copy_(mask, non_blocking=True)
    torch.cuda.current_stream(rays.device).synchronize()
    return h_rays.numpy(), h_mask.numpy()


def _row_dicts(names, rays, mask, n_found):
    """Host rows-layout arrays -> (found, lost) history dicts: zero-copy views, found rays first."""
    n = rays.shape[2]
    found, lost = {}, {}
    for e, name in enumerate(names):
        for box, lo, hi in ((found, 0, n_found), (lost, n_found, n)):
            v = rows_views(rays, mask, e, lo, hi)
            v['mask'] = v['mask'].view(np.bool_)
            box[name] = {key: v[key] for key in RAY_KEYS}
    return found, lost


def gather_rows(torch, rays, mask, n_found):
    """
    Variable-length gather of rows-layout histories onto rank 0 without pickling: one all_gather of the
    (found, lost) counts, then each rank sends its two buffers as they are (NCCL send / recv on the GPUs, gloo in
    the CPU tests) and rank 0 copies the row ranges into place -- found rays of all ranks first (rank order =
    ascending global ray id), then the lost samples.  Returns (rays, mask, n_found_total); other ranks keep their own.
    """
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    n_elem, n = rays.shape[0], rays.shape[2]
    mine = torch.tensor([n_found, n - n_found], dtype=torch.int64, device=rays.device)
    table = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(table, mine)
    table = [[int(v) for v in t.cpu()] for t in table]
    if rank != 0:
        if n:
            for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, rays.contiguous(), 0),
                                               dist.P2POp(dist.isend, mask.contiguous(), 0)]):
                req.wait()
        return rays, mask, n_found
    tot_found = sum(t[0] for t in table)
    tot = tot_found + sum(t[1] for t in table)
    out_rays = torch.empty((n_elem, 7, max(tot, 1)), dtype=rays.dtype, device=rays.device)
    out_mask = torch.empty((n_elem, max(tot, 1)), dtype=mask.dtype, device=rays.device)
    # all receives are posted at once (one batched NCCL group), then the row ranges are copied into place
    parts, ops = {0: (rays, mask)}, []
    for r in range(1, world):
        nf, nl = table[r]
        if nf + nl:
            parts[r] = (torch.empty((n_elem, 7, nf + nl), dtype=rays.dtype, device=rays.device),
                        torch.empty((n_elem, nf + nl), dtype=mask.dtype, device=rays.device))
            ops += [dist.P2POp(dist.irecv, parts[r][0], r), dist.P2POp(dist.irecv, parts[r][1], r)]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    f_at, l_at = 0, tot_found
    for r in range(world):
        nf, nl = table[r]
        if r not in parts:
            continue
        part_rays, part_mask = parts[r]
        for e in range(n_elem):
            if nf:
                src, dst = rows_views(part_rays, part_mask, e, 0, nf), rows_views(out_rays[:, :, :tot], out_mask[:, :tot], e, f_at, f_at + nf)
                for key in RAY_KEYS:
                    dst[key].copy_(src[key])
            if nl:
                src, dst = rows_views(part_rays, part_mask, e, nf, nf + nl), rows_views(out_rays[:, :, :tot], out_mask[:, :tot], e, l_at, l_at + nl)
                for key in RAY_KEYS:
                    dst[key].copy_(src[key])
        f_at += nf
        l_at += nl
    return out_rays[:, :, :tot], out_mask[:, :tot], tot_found


def run_iteration(tracer, stream_id, keep_history=True, keep_images=True, keep_meta=True, max_lost=1000):
    """
    ``_raytrace_iter`` + ``_sort_raytrace`` (xicsrt_raytrace.py:178-278) on the device;
    returns the sorted single-iteration dict.  On ranks > 0 of a multi-GPU run the
    reduced meta / images are still returned; histories hold the rank's own rays.
    """
    out = _skeleton(tracer.config)
    tracer.begin_iteration(stream_id)
    names = tracer.layout.element_names
    if keep_history:
        lost_quota = max_lost if tracer.world == 1 else max(1, max_lost // tracer.world)
        found, lost = tracer.select_ids(stream_id, lost_quota, keep_images)
    else:
        tracer.trace(stream_id, keep_images)
    tracer.allreduce()
    meta, image = tracer.counts_and_images(keep_images)
    if keep_meta:
        out['total']['meta'] = {name: {'num_out': meta[name]} for name in names}
    if keep_images:
        out['total']['image'] = image
    if keep_history:
        # the replay writes the reference's row arrays directly (found rays first, then the lost sample); one
        # device -> host copy through pinned memory, the dict entries are views of it
        n_found = int(found.numel())
        ids = tracer.torch.cat([found, lost])
        rays, mask = tracer.history(stream_id, ids, rows=True)
        if tracer.world > 1:
            rays, mask, n_found = gather_rows(tracer.torch, rays, mask, n_found)
        h_rays, h_mask = to_host(tracer.torch, rays, mask)
        out['found']['history'], out['lost']['history'] = _row_dicts(names, h_rays, h_mask, n_found)
    return out


def merge_histories(parts, names):
    """Concatenate per-rank {element: rays} dicts in rank order (= ascending global ray id for 'found')."""
    return {name: {key: np.concatenate([p[name][key] for p in parts]) for key in RAY_KEYS} for name in names}


def run_iterations_fused(tracer, num_iter, keep_images=True):
    """
    History off: what ``combine_raytrace`` makes of the iterations is the sum of their counters and
    images (xicsrt_raytrace.py:327-356), and those are already accumulated on the device.  All
    iterations are therefore enqueued back to back into the same packed buffer -- no host
    round trip between them -- and reduced and read back once.  (A 1e6-ray iteration is 25 us of
    kernel time; a per-iteration device->host read would cost ten times that.)
    """
    out = _skeleton(tracer.config)
    names = tracer.layout.element_names
    for it in range(num_iter):
        tracer.begin_iteration(it)
        tracer.trace(it, keep_images, zero=(it == 0))
    tracer.allreduce()
    meta, image = tracer.counts_and_images(keep_images)
    out['total']['meta'] = {name: {'num_out': meta[name]} for name in names}
    if keep_images:
        out['total']['image'] = image
    return out


def combine_raytrace(input_list, keep_images=True, components=None):
    """Sum meta and images, concatenate histories (xicsrt_raytrace.py:281-393)."""
    out = _skeleton(input_list[0]['config'])
    names = list(input_list[0]['total']['meta'].keys()) if components is None else list(components)

    for name in names:
        out['total']['meta'][name] = {}
        for key in input_list[0]['total']['meta'][name]:
            out['total']['meta'][name][key] = 0
            for part in input_list:
                out['total']['meta'][name][key] += part['total']['meta'][name][key]

    if keep_images:
        for name in names:
            if name not in input_list[0]['total']['image']:
                continue
            first = input_list[0]['total']['image'][name]
            if first is None:
                out['total']['image'][name] = None
            elif all(part['total']['image'][name].shape == first.shape for part in input_list):
                acc = np.zeros(first.shape)
                for part in input_list:
                    acc += part['total']['image'][name]
                out['total']['image'][name] = acc
            else:
                log.warning('Image dimensions do not match. Cannot combine images.')
                out['total']['image'][name] = None

    if len(input_list[0]['found']['history']) > 0:
        for kind in ('found', 'lost'):
            for name in names:
                if len(input_list) == 1:       # nothing to concatenate: hand the arrays on as they are
                    out[kind]['history'][name] = {key: input_list[0][kind]['history'][name][key] for key in RAY_KEYS}
                else:
                    out[kind]['history'][name] = {
                        key: np.concatenate([part[kind]['history'][name][key] for part in input_list])
                        for key in RAY_KEYS}
    return out


def print_raytrace(results):
    """xicsrt_raytrace.py:414-430."""
    names = list(results['total']['meta'].keys())
    num_source = results['total']['meta'][names[0]]['num_out']
    num_detector = results['total']['meta'][names[-1]]['num_out']
    print('')
    print('Rays Generated: {:6.3e}'.format(num_source))
    print('Rays Detected:  {:6.3e}'.format(num_detector))
    print('Efficiency:     {:6.3e} ± {:3.1e} ({:7.5f}%)'.format(
        num_detector / num_source, np.sqrt(num_detector) / num_source, num_detector / num_source * 100))
    print('')


def _dist_info():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def _resolve_seed(seed, world):
    """``random_seed=None`` means fresh entropy (np.random.seed(None)); all ranks must agree on it."""
    if seed is None:
        seed = int.from_bytes(os.urandom(8), 'little') >> 1
        if world > 1:
            import torch.distributed as dist
            box = [seed]
            dist.broadcast_object_list(box, src=0)
            seed = box[0]
    return int(seed)


# Prepared scenes of the last few calls (elements prepared, tables on the device, buffers allocated), keyed by the content
# of everything they depend on.  A user loop over seeds / runs / repeated calls with one set of elements then pays the
# element preparation and the upload once (0.4 - 0.5 ms of host time per call for the spectrometer, more for a mesh).
# XRT_SCENE_CACHE = number of scenes kept (default 2; 0 = off).
_SCENES = collections.OrderedDict()


def _scene_key(config, rank, world):
    """Content key of a prepared scene, or None where a scene must not be reused."""
    import pickle
    if int(os.environ.get('XRT_SCENE_CACHE', '2')) <= 0:
        return None
    for c in config['sources'].values():
        # plasma sources and Poisson ray counts draw from the run's seed while the scene is prepared
        if str(c.get('class_name', '')).startswith('XicsrtPlasma') or c.get('use_poisson'):
            return None
    for section in ('sources', 'optics', 'filters'):
        for c in (config.get(section) or {}).values():
            # tables read from files (rocking curves, profiles, meshes) may change on disk between calls
            if any('file' in str(k) and v is not None and v != '' for k, v in c.items() if k != 'rocking_type') \
                    or c.get('rocking_type') == 'file':
                return None
    # the library reads its XRT_* switches (tests, measurement knobs) when a scene is created and when it is launched
    raw = getattr(os.environ, '_data', None)
    if isinstance(raw, dict):       # CPython: undecoded view, 6x cheaper than os.environ.items()
        env = tuple(sorted(kv for kv in raw.items() if kv[0][:4] in (b'XRT_', 'XRT_')))
    else:
        env = tuple(sorted((k, v) for k, v in os.environ.items() if k.startswith('XRT_')))
    try:
        torch = _torch()
        body = pickle.dumps((config['sources'], config['optics'], config.get('filters'), config.get('scenario'),
                             config['general'].get('strict_config_check'), rank, world, torch.cuda.current_device(), env),
                            protocol=4)
    except Exception:       # noqa: BLE001  (a user object that does not pickle: prepare the scene afresh)
        return None
    return body


def _acquire_tracer(config, seed, rank, world):
    key = _scene_key(config, rank, world)
    tracer = _SCENES.pop(key, None) if key is not None else None
    if tracer is not None:
        tracer.rebind(config, seed)
        return tracer, key
    tracer = Tracer(config, seed, rank=rank, world=world)
    if key is not None:
        tracer._sections = {k: copy.deepcopy(tracer.config[k]) for k in ('sources', 'optics', 'filters') if k in tracer.config}
    return tracer, key


def _release_tracer(tracer, key, failed=False):
    if key is None or failed:
        tracer.close()
        return
    _SCENES[key] = tracer
    while len(_SCENES) > int(os.environ.get('XRT_SCENE_CACHE', '2')):
        _SCENES.popitem(last=False)[1].close()


def clear_scene_cache():
    """Free the cached scenes (device tables and buffers)."""
    while _SCENES:
        try:
            _SCENES.popitem()[1].close()
        except Exception:       # noqa: BLE001  (interpreter shutdown: the library or the context may be gone)
            pass


atexit.register(clear_scene_cache)      # before the CUDA context goes away at interpreter exit


def raytrace_single(config, _internal=False):
    """One run = ``number_of_iter`` iterations, combined (xicsrt_raytrace.py:87-175)."""
    config = xconfig.to_numpy(config)
    config = xconfig.get_config(config)
    g = config['general']
    rank, world = _dist_info()

    num_iter = g['number_of_iter']
    max_lost_iter = int(g['history_max_lost'] / num_iter)
    if _internal:
        max_lost_iter = max_lost_iter // g['number_of_runs']
    max_lost_iter = max(int(max_lost_iter), 1)

    tracer, key = _acquire_tracer(config, _resolve_seed(g['random_seed'], world), rank, world)
    failed = True
    try:
        if not g['keep_history'] and g['keep_meta']:
            output = run_iterations_fused(tracer, num_iter, keep_images=g['keep_images'])
        else:
            parts = [run_iteration(tracer, it, keep_history=g['keep_history'], keep_images=g['keep_images'],
                                   keep_meta=g['keep_meta'], max_lost=max_lost_iter)
                     for it in range(num_iter)]
            output = combine_raytrace(parts)
        failed = False
    finally:
        _release_tracer(tracer, key, failed)
    if _internal is False:
        _finish(output, g, single=True)
    # per-run images carry the run suffix and are written for direct calls and for every run inside raytrace()
    # (xicsrt_raytrace.py:168-169: outside the `_internal is False` block)
    if g['save_images'] and _dist_info()[0] == 0:
        from . import io as xio
        xio.save_images(output)
    return output


def _finish(output, g, single=False):
    from . import io as xio
    if g['print_results'] and _dist_info()[0] == 0:
        print_raytrace(output)
    if _dist_info()[0] != 0:
        return
    if g['save_config']:
        xio.save_config(output['config'])
    if g['save_images'] and not single:
        xio.save_images(output)
    if g['save_results']:
        xio.save_results(output)


def raytrace(config):
    """``number_of_runs`` runs with cumulative seeds, combined (xicsrt_raytrace.py:28-84)."""
    config = xconfig.get_config(config)
    g = config['general']
    seed = g['random_seed']
    outputs = []
    for ii in range(g['number_of_runs']):
        config_run = copy.deepcopy(config)
        config_run['general']['output_run_suffix'] = '{:04d}'.format(ii)
        if seed is not None:
            seed += ii       # cumulative 0, 1, 3, 6, ... as in the reference (:61-63)
        config_run['general']['random_seed'] = seed
        outputs.append(raytrace_single(config_run, _internal=True))
    output = combine_raytrace(outputs)
    output['config']['general']['output_run_suffix'] = g['output_run_suffix']
    output['config']['general']['random_seed'] = g['random_seed']
    _finish(output, g)
    return output


def raytrace_mp(config, processes=None):
    """
    The reference's multiprocessing entry (xicsrt_multiprocessing.py:12-81) splits *runs*
    over host processes.  Here the parallel resource is the GPU (and, under torchrun, the
    GPUs of the box: rays of every iteration are sharded over ranks inside
    :func:`raytrace_single`), so this is :func:`raytrace`.  ``processes`` has no meaning for a GPU
    run: it is accepted for drop-in compatibility and a value other than None / 1 is logged, not
    silently dropped.
    """
    if processes not in (None, 1):
        log.warning('raytrace_mp: processes=%s ignored -- runs execute on the GPU(s) of this process group', processes)
    return raytrace(config)
