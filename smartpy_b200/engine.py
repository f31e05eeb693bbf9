"""Batch engine: device-resident forcing/observations + one kernel launch per batch of members.

This is the thin host layer between the drop-in API (``smart.py``, ``montecarlo``) and the C
ABI (``include/smart_b200.h``).  PyTorch is used only as plumbing: device memory, pinned
staging buffers and the CUDA stream handle that is passed to the library.

Replaces, for a whole batch at once, the per-sample call chain of the reference:
``MonteCarlo.simulation`` -> ``SMART.simulate`` -> ``structure.run`` -> ``run_all_steps``
(smartpy/montecarlo/montecarlo.py:179-186, smartpy/smart.py:204-210,
smartpy/structure.py:30-197) and ``MonteCarlo.objectivefunction`` (montecarlo.py:193-209).
"""
import ctypes
import math

import numpy as np

from . import _native

SCORE_NAMES = ['NSE', 'KGE', 'KGEc', 'KGEa', 'KGEb', 'PBias', 'RMSE', 'GW']

# engine-level flag (not passed to the library): always hand the kernel one forcing row per step
FLAG_NO_BLOCK_MODE = 0x10000
FLAG_NO_REORDER = 0x20000     # engine-level: keep the members in sample order inside the launch
FLAG_NO_RELAY = 0x40000       # engine-level: no launch scratch for the relay (one CTA walks a group's whole timeline)
STAGE_CHUNKS = 8              # run_host: row chunks of the parameter upload (host copy of one overlaps the DMA of the last)
STAGE_CHUNK_ROWS = 100000     # at least 8 MB of parameters per chunk
REORDER_MIN_MEMBERS = 4096    # below this the sort costs more than the divergence it removes


def _torch():
    import torch
    return torch


def _require_cuda(device=None):
    torch = _torch()
    if not torch.cuda.is_available():
        raise RuntimeError("smartpy_b200 needs a CUDA device (B200, sm_100a): there is no CPU fallback.")
    return torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)


def report_type_of(report):
    """structure.py:65-70."""
    if report == 'summary':
        return _native.REPORT_SUMMARY
    if report == 'raw':
        return _native.REPORT_RAW
    raise Exception('Reporting type \'{}\' unknown.'.format(report))


class BatchEngine(object):
    """Holds one forcing set on the device and runs batches of members through the kernel.

    rain, peva : array [T] (one catchment) or [T, C] (C catchments, read as [t][catchment])
    area_m2    : float or array [C]
    obs        : optional array [n_report] or [n_report, C], NaN = missing
    """

    def __init__(self, rain, peva, area_m2, delta_sec, report_gap, obs=None, extra=None,
                 warm_up_steps=0, report='summary', gw_constraint=None,
                 members_per_catchment=None, precision='f64', device=None, flags=0,
                 forcing_repeat=1):
        torch = _torch()
        self.lib = _native.load()
        self.device = _require_cuda(device)
        if precision not in ('f64', 'f32'):
            raise ValueError("precision must be 'f64' or 'f32'")
        self.precision = precision
        self.flags = int(flags)
        self.report_type = report_type_of(report)
        self.report_gap = int(report_gap)
        self.delta_sec = float(delta_sec)
        self.warm_up_steps = int(warm_up_steps)
        self.extra = extra
        self.gw_constraint = gw_constraint
        self._order_work = {}     # per stream: (workspace, order slots) of smart_member_order
        self._launch_work = {}    # per stream: scratch of smart_batch_run_* (smart_batch_workspace_bytes)
        self._staging = {}        # persistent pinned host buffers of run_host(), by name

        rain = torch.as_tensor(np.ascontiguousarray(rain, dtype=np.float64) if not torch.is_tensor(rain) else rain)
        peva = torch.as_tensor(np.ascontiguousarray(peva, dtype=np.float64) if not torch.is_tensor(peva) else peva)
        if rain.shape != peva.shape:
            raise ValueError("rain and peva must have the same shape")
        self.n_catchments = 1 if rain.dim() == 1 else int(rain.shape[1])
        self.members_per_catchment = int(members_per_catchment) if members_per_catchment else 0
        rain = rain.to(self.device, torch.float64).contiguous()
        peva = peva.to(self.device, torch.float64).contiguous()
        # Block-constant forcing.  forcing_repeat = k > 1: rain/peva are totals over k simulation
        # steps (e.g. daily values for an hourly model) and the equal split of
        # timeframe.py:167-186 (value / k, one IEEE division) happens on the device.  Per-step
        # series that are constant inside aligned blocks of report_gap steps -- what that split
        # produces -- are detected and folded back to one row per block.  Either way the kernel
        # then walks a dry block in closed form (smart_block_fast) when the run qualifies.
        repeat = int(forcing_repeat)
        gap = self.report_gap
        if repeat > 1:
            # (not `tensor / k`: torch multiplies by the reciprocal, which is not the reference's division)
            rows_rain, rows_peva = self._stamp(rain, 1, float(repeat)), self._stamp(peva, 1, float(repeat))
        else:
            rows_rain, rows_peva = rain, peva
            n = int(rain.shape[0])
            # candidate block lengths: the reporting step, and one day of simulation steps when the
            # reporting step divides it (hourly reports of a model forced with daily totals)
            day = 86400.0 / self.delta_sec
            candidates = [gap] + ([int(day)] if day == int(day) and int(day) % gap == 0 else [])
            for k in candidates:
                if k > 1 and n % k == 0 and n >= k and not (self.flags & FLAG_NO_BLOCK_MODE):
                    folded = self._fold(rain, k), self._fold(peva, k)
                    if folded[0] is not None and folded[1] is not None:
                        rows_rain, rows_peva, repeat = folded[0], folded[1], k
                        break
        self._rows = (rows_rain, rows_peva)
        self._repeat = repeat
        self._hourly = None if repeat > 1 else (rows_rain, rows_peva)
        self.n_steps = int(rows_rain.shape[0]) * repeat
        area = np.atleast_1d(np.asarray(area_m2, dtype=np.float64))
        if area.shape != (self.n_catchments,):
            raise ValueError("area_m2 must have one value per catchment")
        self.area = torch.from_numpy(area).to(self.device)

        if self.report_type == _native.REPORT_SUMMARY:
            self.n_report = self.n_steps // self.report_gap
        else:
            self.n_report = -(-self.n_steps // self.report_gap)

        self.obs = None
        self.obs_stats = None
        if obs is not None:
            obs_t = torch.as_tensor(np.ascontiguousarray(obs, dtype=np.float64) if not torch.is_tensor(obs) else obs)
            if obs_t.shape[0] != self.n_report:
                raise ValueError("obs must have n_report = {} rows".format(self.n_report))
            self.obs = obs_t.to(self.device, torch.float64).contiguous()
            self.obs_stats = torch.empty((self.n_catchments, _native.OBS_STATS), dtype=torch.float64,
                                         device=self.device)
            with torch.cuda.device(self.device):
                rc = self.lib.smart_obs_stats(self.obs.data_ptr(), self.n_report, self.n_catchments,
                                              self.obs_stats.data_ptr(),
                                              torch.cuda.current_stream(self.device).cuda_stream)
            _native.check(rc)

    def _stamp(self, rows, repeat, divisor):
        """out[i * repeat + k] = rows[i] / divisor on the device (smart_stamp): the equal split of
        timeframe.py:180-183 with divisor = repeat, a plain expansion with divisor = 1."""
        torch = _torch()
        shape = (rows.shape[0] * repeat,) + tuple(rows.shape[1:])
        out = torch.empty(shape, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.smart_stamp(rows.data_ptr(), rows.shape[0], self.n_catchments, repeat, float(divisor),
                                      out.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream)
        _native.check(rc)
        return out

    def _expand(self, rows, repeat):
        return self._stamp(rows, repeat, 1.0)

    def _fold(self, series, k):
        """One row per aligned block of k rows if the series is constant inside every block (bit for
        bit, smart_fold_blocks), else None."""
        torch = _torch()
        out = torch.empty((series.shape[0] // k,) + tuple(series.shape[1:]), dtype=torch.float64, device=self.device)
        flag = torch.empty((1,), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.smart_fold_blocks(series.data_ptr(), series.shape[0], self.n_catchments, k, out.data_ptr(),
                                            flag.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream)
        _native.check(rc)
        return out if int(flag.item()) == 1 else None

    @property
    def kernel_launches(self):
        """Kernels the library has launched so far (its own counter, smart_launch_count)."""
        return int(self.lib.smart_launch_count())

    @property
    def rain(self):
        """Per-step rainfall on the device ([T] or [T, C])."""
        return self._per_step()[0]

    @property
    def peva(self):
        return self._per_step()[1]

    def _per_step(self):
        if self._hourly is None:
            self._hourly = (self._expand(self._rows[0], self._repeat), self._expand(self._rows[1], self._repeat))
        return self._hourly

    def _block_mode(self, initial_state, last_state):
        """Does this run qualify for one forcing row per reporting step (include/smart_b200.h)?"""
        k = self._repeat
        return (k > 1 and k % self.report_gap == 0 and self.warm_up_steps % k == 0 and initial_state is None and
                not last_state and not (self.flags & FLAG_NO_BLOCK_MODE))

    # ------------------------------------------------------------------ descriptor
    def _desc(self, n_members):
        d = _native.BatchDesc()
        d.n_members = n_members
        d.n_steps = self.n_steps
        d.n_warmup = self.warm_up_steps
        d.n_catchments = self.n_catchments
        d.members_per_catchment = self.members_per_catchment if self.n_catchments > 1 else 1
        d.report_gap = self.report_gap
        d.report_type = self.report_type
        d.flags = self.flags & 0xffff
        d.dt_sec = self.delta_sec
        d.area_m2 = self.area.data_ptr()
        if self.extra:   # truthiness, as structure.py:100
            d.has_extra = 1
            d.aar = float(self.extra['aar'])
            d.ro_ratio = float(self.extra['r-o_ratio'])
            for k in range(5):
                d.ro_split[k] = float(self.extra['r-o_split'][k])
        gwc = self.gw_constraint
        d.gw_constraint = float(gwc) if gwc else float('nan')   # truthiness, as montecarlo.py:71-74
        return d

    # ------------------------------------------------------------------ run
    def run(self, params, discharge=False, scores=None, gw=True, last_state=False,
            initial_state=None, best=None, out=None):
        """Run every row of params[N, 10] (names order T,C,H,D,S,Z,SK,FK,GK,RK).

        params may be a numpy array (copied host->device through pinned memory) or a CUDA
        float64 tensor (used in place).  Returns a dict of device tensors:
          'discharge' [n_report, N] (m3/s), 'scores' [N, 8], 'gw' [N], 'last_state' [N, 19],
          'best' (score, index) -- only those requested.  Stream-ordered; nothing synchronises.
        """
        torch = _torch()
        dev = self.device
        if scores is None:
            scores = self.obs is not None
        if torch.is_tensor(params):
            p_dev = params.to(dev, torch.float64).contiguous()
        else:
            p_host = np.ascontiguousarray(params, dtype=np.float64)
            if p_host.ndim == 1:
                p_host = p_host[None, :]
            pinned = self._pinned('run_params', p_host.shape, torch.float64)
            torch.cuda.current_stream(dev).synchronize()     # the previous copy out of this buffer is done
            pinned.copy_(torch.from_numpy(p_host))
            p_dev = pinned.to(dev, non_blocking=True)
        if p_dev.dim() != 2 or p_dev.shape[1] != _native.N_PARAMS:
            raise ValueError("params must be [N, 10]")
        n = int(p_dev.shape[0])
        d = self._desc(n)
        d.params = p_dev.data_ptr()
        if self._block_mode(initial_state, last_state):
            d.rain, d.peva = self._rows[0].data_ptr(), self._rows[1].data_ptr()
            d.forcing_repeat = self._repeat
        else:
            d.rain, d.peva = self.rain.data_ptr(), self.peva.data_ptr()
        keep = [p_dev]
        res = {}
        out = out or {}
        qdtype = torch.float64 if self.precision == 'f64' else torch.float32
        if discharge:
            q = out.get('discharge')
            if q is None:
                q = torch.empty((self.n_report, n), dtype=qdtype, device=dev)
            d.discharge = q.data_ptr()
            d.ld_discharge = q.stride(0)
            res['discharge'] = q
        if (scores or best) and self.obs is None:
            raise Exception("scores / best member requested but the engine has no observations")
        if scores or best:
            d.obs = self.obs.data_ptr()
            d.obs_stats = self.obs_stats.data_ptr()
        # out['block']: a caller-owned [n, 9] tensor; scores go to its columns 0..7 and gw to column 8
        # (the row block a sharded run all-gathers), written by the kernel itself
        blk = out.get('block')
        if blk is not None and (blk.shape != (n, _native.N_SCORES + 1) or blk.dtype != torch.float64 or
                                not blk.is_contiguous() or blk.device != dev):
            raise ValueError("out['block'] must be a contiguous float64 [N, 9] tensor on the engine's device")
        if scores:
            sc = blk[:, :_native.N_SCORES] if blk is not None else out.get('scores')
            if sc is None:
                sc = torch.empty((n, _native.N_SCORES), dtype=torch.float64, device=dev)
            d.scores = sc.data_ptr()
            d.ld_scores = sc.stride(0) if n > 1 else _native.N_SCORES
            res['scores'] = sc
        if gw:
            g = blk[:, _native.N_SCORES] if blk is not None else out.get('gw')
            if g is None:
                g = torch.empty((n,), dtype=torch.float64, device=dev)
            d.gw = g.data_ptr()
            d.ld_gw = g.stride(0) if n > 1 else 1
            res['gw'] = g
        if last_state:
            ls = torch.empty((n, _native.N_VARS), dtype=torch.float64, device=dev)
            d.last_state = ls.data_ptr()
            res['last_state'] = ls
        if initial_state is not None:
            ini = torch.as_tensor(np.ascontiguousarray(initial_state, dtype=np.float64)
                                  if not torch.is_tensor(initial_state) else initial_state)
            ini = ini.reshape(-1, _native.N_VARS).to(dev, torch.float64).contiguous()
            if ini.shape[0] != n:
                raise ValueError("initial_state must be [N, 19]")
            d.initial_state = ini.data_ptr()
            keep.append(ini)
        fn = self.lib.smart_batch_run_f64 if self.precision == 'f64' else self.lib.smart_batch_run_f32
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev)
            if (self.n_catchments == 1 and n >= REORDER_MIN_MEMBERS and not discharge and not last_state
                    and initial_state is None and not (self.flags & FLAG_NO_REORDER)):
                order = self._member_order(p_dev, n, stream)
                d.member_order = order.data_ptr()
                d.member_order_len = order.numel()
            if best:
                column, sign = best
                d.best_column = SCORE_NAMES.index(column) if isinstance(column, str) else int(column)
                d.best_sign = 1 if sign > 0 else -1
                bs = torch.empty(1, dtype=torch.float64, device=dev)
                bi = torch.empty(1, dtype=torch.int64, device=dev)
                d.best_score = bs.data_ptr()
                d.best_index = bi.data_ptr()
                res['best'] = (bs, bi)
            # scratch of the launch: per-CTA winners of the best-member search and, for batches of a
            # few waves, the state the relay parks between two segments of the timeline (smart_b200.h)
            ws_bytes = int(self.lib.smart_batch_workspace_bytes(ctypes.byref(d)))
            if ws_bytes > 0 and (best or not (self.flags & FLAG_NO_RELAY)):
                ws = self._workspace(ws_bytes, stream)
                d.workspace = ws.data_ptr()
                d.workspace_bytes = 0 if (self.flags & FLAG_NO_RELAY) else ws.numel()
            rc = fn(ctypes.byref(d), stream.cuda_stream)
        _native.check(rc)
        for t in keep:   # keep inputs alive until the stream has consumed them
            t.record_stream(stream)
        return res

    def _workspace(self, nbytes, stream):
        """Launch scratch, kept per stream (runs on one stream are ordered) and grown on demand."""
        torch = _torch()
        ws = self._launch_work.get(stream.cuda_stream)
        if ws is None or ws.numel() < nbytes:
            ws = self._launch_work[stream.cuda_stream] = torch.empty((nbytes,), dtype=torch.uint8, device=self.device)
        return ws

    def _member_order(self, p_dev, n, stream):
        """Slots of the launch (smart_member_order): members grouped so that the lanes of a warp take
        the same branches -- inside the fast form's domain first (padded to a CTA boundary), the
        others after; each group by slice of T, then by S * Z.  Results do not change by a bit: a
        member's arithmetic never depends on its neighbours.  Workspace and slots are kept per
        stream (two runs on one stream are ordered; runs on different streams do not share them)."""
        torch = _torch()
        slots = int(self.lib.smart_member_order_len(n))
        nbytes = int(self.lib.smart_member_order_workspace_bytes(n))
        work, order = self._order_work.get(stream.cuda_stream, (None, None))
        if work is None or work.numel() < nbytes or order.numel() != slots:
            work = torch.empty((nbytes,), dtype=torch.uint8, device=self.device)
            order = torch.empty((slots,), dtype=torch.int64, device=self.device)
            self._order_work[stream.cuda_stream] = (work, order)
        _native.check(self.lib.smart_member_order(p_dev.data_ptr(), n, self.delta_sec, order.data_ptr(),
                                                  work.data_ptr(), stream.cuda_stream))
        return order

    # ------------------------------------------------------------------ host buffers in, host buffers out
    def _pinned(self, name, shape, dtype):
        torch = _torch()
        buf = self._staging.get(name)
        if buf is None or buf.shape != tuple(shape) or buf.dtype != dtype:
            buf = self._staging[name] = torch.empty(tuple(shape), dtype=dtype).pin_memory()
        return buf

    def run_host(self, params, discharge=False, gw=True, copy=False):
        """params: numpy [N, 10] on the host, or a page-locked tensor from pinned_rows() (no staging
        copy) -> dict of numpy arrays on the host ('scores' [N, 8] when
        the engine has observations, 'gw' [N], 'discharge' [n_report, N] on request).  The copies go
        through pinned staging buffers the engine keeps from call to call (no pin_memory() or
        allocation on the steady path) and the call returns when the results are in host memory.
        The returned arrays are VIEWS of that staging memory: they are overwritten by the next
        run_host() of this engine; pass copy=True (or copy them) to keep them longer."""
        torch = _torch()
        dev = self.device
        if torch.is_tensor(params):
            p_host = params if params.dim() == 2 else params[None, :]
        else:
            p_host = np.ascontiguousarray(params, dtype=np.float64)
            if p_host.ndim == 1:
                p_host = p_host[None, :]
        n = p_host.shape[0]
        scored = self.obs is not None
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev)
            p_dev = self.stage_params(p_host)
            blk = self._device_buffer('block', (n, _native.N_SCORES + 1))
            res = self.run(p_dev, discharge=discharge, scores=scored, gw=True, out={'block': blk})
            blk_pin = self._pinned('block', blk.shape, torch.float64)
            blk_pin.copy_(blk, non_blocking=True)
            out = {}
            if discharge:
                q_pin = self._pinned('discharge', res['discharge'].shape, res['discharge'].dtype)
                q_pin.copy_(res['discharge'], non_blocking=True)
            stream.synchronize()
        table = blk_pin.numpy()
        if scored:
            out['scores'] = table[:, :_native.N_SCORES]
        if gw:
            out['gw'] = table[:, _native.N_SCORES]
        if discharge:
            out['discharge'] = q_pin.numpy()
        if copy:
            out = {k: v.copy() for k, v in out.items()}
        return out

    @staticmethod
    def pinned_rows(n_members):
        """A page-locked [N, 10] float64 host tensor for the caller to fill (`.numpy()` is a view of it)
        and hand to run_host(): the upload is then one DMA straight out of it, with no staging copy."""
        torch = _torch()
        return torch.empty((int(n_members), _native.N_PARAMS), dtype=torch.float64).pin_memory()

    def stage_params(self, p_host):
        """Host rows [N, 10] -> the engine's device buffer on the current stream.  A page-locked float64
        tensor (pinned_rows()) is copied by the DMA engine directly; a numpy array goes through the
        engine's pinned staging buffer, large ones in row chunks so that the DMA of one chunk runs while
        the host copies the next (C4a, 80 MB of parameters per 30 ms run: e2e 2.55e11 -> 2.61e11;
        batches of less than 16 MB go over in one piece)."""
        torch = _torch()
        if torch.is_tensor(p_host):
            if not (p_host.device.type == 'cpu' and p_host.is_pinned() and p_host.dtype == torch.float64
                    and p_host.is_contiguous()):
                p_host = p_host.detach().cpu().contiguous().to(torch.float64).numpy()
            else:
                p_dev = self._device_buffer('params', tuple(p_host.shape))
                p_dev.copy_(p_host, non_blocking=True)
                return p_dev
        n = p_host.shape[0]
        p_pin = self._pinned('params', p_host.shape, torch.float64)
        p_dev = self._device_buffer('params', p_host.shape)
        src = torch.from_numpy(p_host)
        chunks = max(1, min(STAGE_CHUNKS, n // STAGE_CHUNK_ROWS))   # (each copy call costs ~20 us: no chunks below 8 MB)
        step = -(-n // chunks)
        for lo in range(0, n, step):
            hi = min(n, lo + step)
            p_pin[lo:hi].copy_(src[lo:hi])
            p_dev[lo:hi].copy_(p_pin[lo:hi], non_blocking=True)
        return p_dev

    def _device_buffer(self, name, shape):
        torch = _torch()
        buf = self._staging.get('dev_' + name)
        if buf is None or buf.shape != tuple(shape):
            buf = self._staging['dev_' + name] = torch.empty(tuple(shape), dtype=torch.float64, device=self.device)
        return buf

    # steps one launch of n members walks through (warm-up + main), for throughput accounting
    def member_steps(self, n_members):
        return int(n_members) * (self.n_steps + self.warm_up_steps)


def warm_up_length(warm_up_days, delta_sec):
    """structure.py:88."""
    return int(warm_up_days * 86400 / delta_sec)


def fma_peak(precision=64, blocks=None, threads=256, iters=1 << 16, repeats=5, device=None):
    """Measured FMA-pipe peak of the visible GPU in FMA instructions per second (per thread-op).

    Runs the library's register-resident FMA probe and times it with CUDA events on the
    launching stream.  Returns (fma_per_second, ms).
    """
    torch = _torch()
    dev = _require_cuda(device)
    lib = _native.load()
    sm = torch.cuda.get_device_properties(dev).multi_processor_count
    blocks = blocks or sm * 8
    out = torch.empty(blocks * threads, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev)
    best = math.inf
    with torch.cuda.device(dev):
        for i in range(repeats + 1):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            _native.check(lib.smart_fma_peak_probe(precision, blocks, threads, iters, out.data_ptr(),
                                                   stream.cuda_stream))
            e1.record(stream)
            e1.synchronize()
            if i:
                best = min(best, e0.elapsed_time(e1))
    fmas = blocks * threads * 8 * iters
    return fmas / (best * 1e-3), best


def score_series(simulation, observation, device=None):
    """NSE, KGE, KGEc, KGEa, KGEb, PBias, RMSE of one simulated series against observations
    with NaN = missing (montecarlo.py:193-203), computed on the device by
    ``smart_score_discharge``.  Returns a list of 7 floats."""
    torch = _torch()
    dev = _require_cuda(device)
    lib = _native.load()
    sim = torch.as_tensor(np.ascontiguousarray(simulation, dtype=np.float64)).to(dev).reshape(-1, 1).contiguous()
    obs = torch.as_tensor(np.ascontiguousarray(observation, dtype=np.float64)).to(dev).reshape(-1, 1).contiguous()
    if sim.shape != obs.shape:
        raise ValueError("simulation and observation must have the same length")
    stats = torch.empty((1, _native.OBS_STATS), dtype=torch.float64, device=dev)
    out = torch.empty((1, _native.N_SCORES), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _native.check(lib.smart_obs_stats(obs.data_ptr(), obs.shape[0], 1, stats.data_ptr(), stream))
        _native.check(lib.smart_score_discharge(sim.data_ptr(), 1, 1, sim.shape[0], obs.data_ptr(),
                                                stats.data_ptr(), 1, 1, 64, out.data_ptr(), stream))
    return out[0, :7].cpu().tolist()
