"""BatchManager of arch=nn (reference data_nn.py:12-168): the training set of the latent-space integrator.

Input: `<code_path>/code<z_num>.npz` written by `Trainer.test_ae` ({x: codes of frames 0..f-2 of every simulation, y: codes
of frames 1..f-1, p: per-frame parameter increments, s: #simulations, f: #frames}) and `<data_path>/args.txt` (num_dof).
Same preprocessing as the reference: y -= x (the network predicts the code INCREMENT), x / std(x), y / std(y), p / std(p),
features = concat(x, p); the first 95 % of the simulations train, the rest test; windows of w_size consecutive frames.
The reference feeds tf.data iterators (`.batch(B).repeat().shuffle(50)`: whole BATCHES pass through a 50-slot shuffle buffer);
here the same buffer discipline runs on a numpy RandomState and the batches are handed out as device tensors.
"""
import os

import numpy as np
import torch


class _ShuffledBatches(object):
    """dataset.batch(B).repeat().shuffle(buffer_size=50): consecutive batches, repeated, drawn through a shuffle buffer"""

    def __init__(self, x, y, batch_size, rng, buffer_size=50):
        self.x, self.y, self.b, self.rng = x, y, batch_size, rng
        self.pos, self.buf = 0, []
        for _ in range(buffer_size):
            self.buf.append(self._next_sequential())

    def _next_sequential(self):
        n = self.x.shape[0]
        if self.pos >= n:
            self.pos = 0
        s = slice(self.pos, min(self.pos + self.b, n))
        self.pos += self.b
        return self.x[s], self.y[s]

    def get(self):
        i = self.rng.randint(len(self.buf))
        out = self.buf[i]
        self.buf[i] = self._next_sequential()
        return out


class _SequentialBatches(object):
    """dataset.batch(B): one pass in order (the test iterators; re-created by init_test_it)"""

    def __init__(self, x, y, batch_size):
        self.x, self.y, self.b, self.pos = x, y, batch_size, 0

    def get(self):
        if self.pos >= self.x.shape[0]:
            raise StopIteration("test iterator exhausted: call init_test_it() (tf.errors.OutOfRangeError in the reference)")
        s = slice(self.pos, self.pos + self.b)
        self.pos += self.b
        return self.x[s], self.y[s]


class BatchManager(object):
    def __init__(self, config, device=None):
        self.rng = np.random.RandomState(config.random_seed)
        self.root = config.data_path
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.args = {}
        with open(os.path.join(self.root, 'args.txt'), 'r') as f:          # data_nn.py:18-25
            for line in f:
                if ': ' in line:
                    k, v = line.rstrip('\n').split(': ', 1)
                    self.args[k] = v
        self.is_3d = config.is_3d
        self.w_num = config.w_size
        self.z_num = config.z_num
        self.dof = int(self.args['num_dof'])
        self.code_path = os.path.join(config.code_path, 'code%d.npz' % self.z_num)
        self.batch_size = config.batch_size

        code = np.load(self.code_path)
        x, y, p = code['x'].copy(), code['y'].copy(), code['p'].copy()
        self.num_scenes, self.num_frames = int(code['s']), int(code['f'])
        self.code_std = np.std(x)                                           # data_nn.py:68-77
        y -= x
        self.out_std = np.std(y)
        self.p_std = np.std(p)
        x /= self.code_std
        y /= self.out_std
        p /= self.p_std
        self.x_train = np.concatenate((x, p), axis=-1)
        self.y_train = y
        self.num_train_scenes = int(self.num_scenes * 0.95)
        self.num_test_scenes = self.num_scenes - self.num_train_scenes
        self.num_train = self.num_train_scenes * (self.num_frames - 1)
        self.num_test = self.x_train.shape[0] - self.num_train
        self.x_test, self.y_test = self.x_train[self.num_train:], self.y_train[self.num_train:]
        self.x_train, self.y_train = self.x_train[:self.num_train], self.y_train[:self.num_train]

        def windows(xs, ys, scenes):                                        # data_nn.py:91-113
            n = scenes * (self.num_frames - self.w_num)
            xw = np.zeros([n, self.w_num, self.z_num + self.dof])
            yw = np.zeros([n, self.w_num, self.z_num])
            k = 0
            for i in range(scenes):
                for j in range(self.num_frames - self.w_num):
                    idx = i * (self.num_frames - 1) + j
                    xw[k] = xs[idx:idx + self.w_num]
                    yw[k] = ys[idx:idx + self.w_num]
                    k += 1
            return xw, yw
        self.x_train_w, self.y_train_w = windows(self.x_train, self.y_train, self.num_train_scenes)
        self.x_test_w, self.y_test_w = windows(self.x_test, self.y_test, self.num_test_scenes)
        self.num_train_w, self.num_test_w = self.x_train_w.shape[0], self.x_test_w.shape[0]
        self.num_samples = self.num_train + self.num_test
        self.train_steps = max(int(self.num_train / self.batch_size + 0.5), 1)          # per epoch
        self.test_steps = max(int(self.num_test / self.batch_size + 0.5), 1)
        self.train_w_steps = max(int(self.num_train_w / self.batch_size + 0.5), 1)
        self.test_w_steps = max(int(self.num_test_w / self.batch_size + 0.5), 1)
        self.epochs_per_step = 1 / self.train_w_steps
        self.c_num = 0
        self.init_it()
        self.init_test_it()

    def init_it(self, sess=None):
        self._train = _ShuffledBatches(self.x_train, self.y_train, self.batch_size, self.rng)
        self._train_w = _ShuffledBatches(self.x_train_w, self.y_train_w, self.batch_size, self.rng)

    def init_test_it(self):
        self._test = _SequentialBatches(self.x_test, self.y_test, self.batch_size)
        self._test_w = _SequentialBatches(self.x_test_w, self.y_test_w, self.batch_size)

    def _dev(self, pair):
        return tuple(torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=self.device) for a in pair)

    def batch(self, is_window=False):
        return self._dev((self._train_w if is_window else self._train).get())

    def test_batch(self, is_window=False):
        return self._dev((self._test_w if is_window else self._test).get())

    def start_thread(self, sess=None):
        pass

    def stop_thread(self):
        pass
