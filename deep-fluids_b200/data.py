"""BatchManager with the attributes the reference's Trainer reads (data.py:16-171): `.batch() -> (x, y)`, `.c_num`,
`.epochs_per_step`, `.q.size()`, `.start_thread/.stop_thread`, `.random_list`, `.denorm`, `.y_num`, `.dof`, `.root`.

The reference's FIFOQueue + GIL-bound npz loader threads (data.py:110-159) are replaced by a synthetic, on-device
tensor source of identical shape and range: params y ~ U[-1,1] [B,c_num], target velocity x = curl of
a smoothed random potential scaled to max|x| = 1 (mirrors x /= x_range, data.py:329).  A pool of `pool` distinct
batches is generated once with the seeded generator and served round-robin.  Real-dataset loading (args.txt,
v/*.npz) is `DatasetBatchManager` below: the reference's file order, ranges and normalisation, loader
threads feeding pinned host batches one step ahead."""
import os

import numpy as np
import torch

from . import kernels as K


class _Queue(object):
    def size(self):
        return 0


def BatchManager(config, device=None, pool=4, rank=0):
    """Factory with the reference's constructor name (data.py:16): the real-dataset loader when `data/<dataset>/args.txt`
    exists and --synthetic is not set, else the synthetic on-device source."""
    root = getattr(config, "data_path", None) or os.path.join(config.data_dir, config.dataset)
    if not getattr(config, "synthetic", False) and os.path.exists(os.path.join(root, "args.txt")):
        return DatasetBatchManager(config, device=device, rank=rank)
    return SyntheticBatchManager(config, device=device, pool=pool, rank=rank)


class SyntheticBatchManager(object):
    def __init__(self, config, device=None, pool=4, rank=0):
        self.config = config
        self.root = getattr(config, "data_path", "")
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.is_3d = config.is_3d
        self.batch_size = config.batch_size
        self.res_x, self.res_y, self.res_z = config.res_x, config.res_y, config.res_z
        self.c_num = 3                       # smoke_pos_size / smoke3_vel_buo: [p0, p1, t]  (data.py:52-60)
        self.y_num = [21, 5, 200] if not self.is_3d else [5, 3, 250]
        self.y_range = [[-1.0, 1.0]] * self.c_num
        self.x_range = 1.0
        self.dof = 2
        self.num_samples = int(getattr(config, "synthetic_samples", 21000))
        self.epochs_per_step = self.batch_size / float(self.num_samples)   # data.py:105
        self.q = _Queue()
        self._pool = []
        g = torch.Generator(device=self.device).manual_seed(int(config.random_seed) + int(rank))
        if self.is_3d:
            sp = [self.res_z, self.res_y, self.res_x]
        else:
            sp = [self.res_y, self.res_x]
        for _ in range(pool):
            self._pool.append(self._make(sp, g))
        self._i = 0

    def _make(self, sp, g):
        nd = len(sp)
        if 'ae' in getattr(self.config, "arch", "de"):
            # AE scenes store the source-position history: y [B, dof, num_frames] (data.py:70-72); only the last
            # frame supervises the latent code (trainer.py:385)
            y = torch.rand(self.batch_size, self.dof, 1, device=self.device, generator=g) * 2 - 1
        else:
            y = torch.rand(self.batch_size, self.c_num, device=self.device, generator=g) * 2 - 1
        pot = torch.randn([self.batch_size] + sp + [1 if nd == 2 else 3], device=self.device, generator=g)
        for _ in range(2):                       # separable box smoothing: band-limits the field
            for ax in range(1, nd + 1):
                pot = (pot + torch.roll(pot, 1, ax) + torch.roll(pot, -1, ax)) / 3.0
        x = K.curl_fwd(pot.contiguous())         # divergence-free by construction (our curl kernel)
        x = x / x.abs().max()
        return x.contiguous(), y.contiguous()

    # ---- the interface Trainer uses
    def batch(self):
        x, y = self._pool[self._i % len(self._pool)]
        self._i += 1
        return x, y

    def start_thread(self, sess=None):
        pass

    def stop_thread(self):
        pass

    def denorm(self, x=None, y=None):
        if x is not None:
            x = x * self.x_range
        return x, y

    def random_list(self, num):
        x, y = self._pool[0]
        return x[:num], None, y[:num]


class HostPrefetcher(object):
    """Pinned host batches -> device, ONE STEP AHEAD, on a dedicated copy stream.

    The reference feeds its graph through loader threads and a FIFOQueue (data.py:124-144): the next batch is already
    queued while `sess.run(g_optim)` works on the current one.  Here the same overlap is explicit: `put(x_host, y_host)`
    starts the H2D copy of the NEXT batch into the free one of two device slots on the copy stream; `get()` makes the
    compute stream wait for the copy of the CURRENT batch and returns its device tensors.  The copy engine runs under the
    step's kernels, so PCIe time (1.8 ms for a 4 x 128^3 x 3 fp32 batch) leaves the critical path."""

    def __init__(self, like_x, like_y, device=None):
        self.device = device if device is not None else like_x.device
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = [(torch.empty_like(like_x, device=self.device), torch.empty_like(like_y, device=self.device)) for _ in range(2)]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.free = [None, None]              # compute-stream event after which a slot may be overwritten
        self.n_put = self.n_get = 0

    def put(self, x_host, y_host):
        assert self.n_put - self.n_get < 2, "HostPrefetcher holds at most two batches"
        k = self.n_put % 2
        if self.free[k] is not None:
            self.stream.wait_event(self.free[k])
        with torch.cuda.stream(self.stream):
            self.slots[k][0].copy_(x_host, non_blocking=True)
            self.slots[k][1].copy_(y_host, non_blocking=True)
            self.ready[k].record(self.stream)
        self.n_put += 1

    def get(self):
        assert self.n_get < self.n_put, "HostPrefetcher.get() without a pending put()"
        k = self.n_get % 2
        torch.cuda.current_stream(self.device).wait_event(self.ready[k])
        self.n_get += 1
        return self.slots[k]

    def release(self):
        """call after the consumer of the last get() has been enqueued on the compute stream"""
        k = (self.n_get - 1) % 2
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.free[k] = ev


def preprocess(file_path, data_type, x_range, y_range):
    """npz {x, y} -> normalised (x, y), as reference data.py:311-333: velocity x /= x_range (density: x*2-1);
    y[i] -> [-1, 1] by its [min, max]."""
    import numpy as np
    with np.load(file_path) as data:
        x = data['x']
        y = np.array(data['y'], dtype=np.float64)
    # The reference divides the fp32 field in place by the float64 range read with np.loadtxt (`x /= x_range`): numpy
    # evaluates that in double and rounds once to fp32; the labels stay double until TensorFlow casts the feed to fp32.
    if data_type[0] == 'd':
        x = (x * 2 - 1).astype(np.float32)
    else:
        x = (x.astype(np.float64) / np.float64(x_range)).astype(np.float32)
    for i, ri in enumerate(y_range):
        y[i] = (y[i] - ri[0]) / (ri[1] - ri[0]) * 2 - 1
    return x, y.astype(np.float32)


class DatasetBatchManager(object):
    """Real mantaflow dataset (reference data.py:16-171): `args.txt` ("key: value" per line), `v/*.npz` with x (velocity
    field) and y (parameters, or [dof, num_frames] history for AE scenes), `v_range.txt` (min/max -> x_range).
    The reference's GIL-bound loader threads + TF FIFOQueue become worker threads that assemble whole batches in pinned
    host memory; `.batch()` hands out device tensors (async H2D on the caller's stream)."""

    def __init__(self, config, device=None, rank=0, prefetch=4):
        import glob
        import numpy as np
        import queue
        self.config = config
        self.root = getattr(config, "data_path", None) or os.path.join(config.data_dir, config.dataset)
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.rng = np.random.RandomState(int(config.random_seed) + int(rank))
        self.args = {}
        with open(os.path.join(self.root, 'args.txt'), 'r') as f:          # data.py:22-29
            for line in f:
                if ': ' in line:
                    arg, val = line.rstrip("\n").split(': ', 1)
                    self.args[arg] = val
        self.is_3d = config.is_3d
        self.data_type = config.data_type
        pat = os.path.join(self.root, self.data_type[0], "*")
        if 'ae' in config.arch:                                             # data.py:32-38: sort by (scene, frame)
            nf = int(self.args['num_frames'])

            def sortf(x):
                n = os.path.basename(x)[:-4].split('_')
                return int(n[0]) * nf + int(n[1])
            self.paths = sorted(glob.glob(pat), key=sortf)
        else:
            self.paths = sorted(glob.glob(pat))
        self.num_samples = len(self.paths)
        assert self.num_samples > 0, "no samples under %s" % pat
        self.batch_size = config.batch_size
        self.epochs_per_step = self.batch_size / float(self.num_samples)   # data.py:52
        self.res_x, self.res_y, self.res_z = config.res_x, config.res_y, config.res_z
        self.depth = (3 if self.is_3d else 2) if self.data_type == 'velocity' else 1
        self.c_num = int(self.args['num_param'])
        self.feature_dim = ([self.res_z] if self.is_3d else []) + [self.res_y, self.res_x, self.depth]
        if 'ae' in config.arch:
            self.dof = int(self.args['num_dof'])
            self.label_dim = [self.dof, int(self.args['num_frames'])]
        else:
            self.dof = 0
            self.label_dim = [self.c_num]
        r = np.loadtxt(os.path.join(self.root, self.data_type[0] + '_range.txt'))   # data.py:87-88
        self.x_range = float(max(abs(r[0]), abs(r[1])))
        self.y_range, self.y_num = [], []
        for i in range(self.c_num):                                          # data.py:92-108
            p_name = self.args['p%d' % i]
            self.y_num.append(int(self.args['num_{}'.format(p_name)]))
            if 'ae' not in config.arch:
                self.y_range.append([float(self.args['min_{}'.format(p_name)]), float(self.args['max_{}'.format(p_name)])])
        if 'ae' in config.arch:
            self.y_range = [[-1, 1] for _ in range(self.label_dim[0])]
        self.num_threads = int(max(1, min(config.num_worker, os.cpu_count() or 1, self.batch_size)))
        self._queue = queue.Queue(maxsize=prefetch)
        self._threads, self._stop = [], False
        self.q = self                       # Trainer reads batch_manager.q.size()

    def size(self):
        return self._queue.qsize()

    # ---- loader threads (data.py:116-159)
    def _worker(self, seed):
        import numpy as np
        rng = np.random.RandomState(seed)
        pin = torch.cuda.is_available()
        while not self._stop:
            xb = torch.empty([self.batch_size] + self.feature_dim, dtype=torch.float32)
            yb = torch.empty([self.batch_size] + self.label_dim, dtype=torch.float32)
            for i in range(self.batch_size):
                x, y = preprocess(self.paths[rng.randint(self.num_samples)], self.data_type, self.x_range, self.y_range)
                xb[i] = torch.from_numpy(np.ascontiguousarray(x)).reshape(self.feature_dim)
                yb[i] = torch.from_numpy(np.ascontiguousarray(y)).reshape(self.label_dim)
            if pin:
                xb, yb = xb.pin_memory(), yb.pin_memory()
            while not self._stop:
                try:
                    self._queue.put((xb, yb), timeout=0.1)
                    break
                except Exception:
                    continue

    def start_thread(self, sess=None):
        import threading
        if self._threads:
            return
        self._stop = False
        for i in range(self.num_threads):
            t = threading.Thread(target=self._worker, args=(int(self.rng.randint(1 << 30)),), daemon=True)
            t.start()
            self._threads.append(t)

    def stop_thread(self):
        self._stop = True
        for t in self._threads:
            t.join(timeout=2.0)
        self._threads = []

    def batch(self):
        """(x, y) of one batch on the device: x [B,(D,)H,W,C] in [-1,1], y [B,c_num] (AE: [B,dof,frames])"""
        self.start_thread()
        xb, yb = self._queue.get()
        return xb.to(self.device, non_blocking=True), yb.to(self.device, non_blocking=True)

    def denorm(self, x=None, y=None):
        if x is not None:
            x = x * self.x_range if self.data_type[0] != 'd' else (x + 1) * 0.5
        if y is not None and 'ae' not in self.config.arch:
            y = y.clone() if hasattr(y, "clone") else y.copy()
            for i, ri in enumerate(self.y_range):
                y[..., i] = (y[..., i] + 1) * 0.5 * (ri[1] - ri[0]) + ri[0]
        return x, y

    def batch_(self, b_num):
        """the whole dataset in FILE ORDER, b_num normalised fields at a time (data.py:173-183; test_ae's latent dump)"""
        assert len(self.paths) % b_num == 0
        xb = []
        for i, path in enumerate(self.paths):
            x, _ = preprocess(path, self.data_type, self.x_range, self.y_range)
            xb.append(x)
            if (i + 1) % b_num == 0:
                yield np.array(xb), []
                xb = []

    def random_list(self, num):
        xs, ys = [], []
        for _ in range(num):
            x, y = preprocess(self.paths[self.rng.randint(self.num_samples)], self.data_type, self.x_range, self.y_range)
            xs.append(torch.from_numpy(x))
            ys.append(torch.from_numpy(y))
        return torch.stack(xs).to(self.device), None, torch.stack(ys).to(self.device)
