"""BatchManager with the attributes the reference's Trainer reads (data.py:16-171): `.batch() -> (x, y)`, `.c_num`,
`.epochs_per_step`, `.q.size()`, `.start_thread/.stop_thread`, `.random_list`, `.denorm`, `.y_num`, `.dof`, `.root`.

The reference's FIFOQueue + GIL-bound npz loader threads (data.py:110-159) are replaced by a synthetic, on-device
tensor source of identical shape and range: params y ~ U[-1,1] [B,c_num], target velocity x = curl of
a smoothed random potential scaled to max|x| = 1 (mirrors x /= x_range, data.py:329).  A pool of `pool` distinct
batches is generated once with the seeded generator and served round-robin.  Real-dataset loading (args.txt,
v/*.npz) is SURVEY 8(f) row N2, not built yet."""
import os

import torch

from . import kernels as K


class _Queue(object):
    def size(self):
        return 0


class BatchManager(object):
    def __init__(self, config, device=None, pool=4, rank=0):
        self.config = config
        self.root = getattr(config, "data_path", "")
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.is_3d = config.is_3d
        self.batch_size = config.batch_size
        self.res_x, self.res_y, self.res_z = config.res_x, config.res_y, config.res_z
        if not getattr(config, "synthetic", False) and os.path.isdir(self.root):
            raise NotImplementedError("real dataset loading (data.py:16-108) is SURVEY 8(f) N2; pass --synthetic=true")
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
