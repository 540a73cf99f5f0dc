"""
Device-side replacement of the reference's per-record input transforms.

The reference builds, per record and on the CPU in numpy (`ecg_transformer/preprocess/ptb_dataset.py:132-149`,
`dataset.py:70-75,87-93`):  `Normalize(mean, std)` (`transform.py:18-35`)  ->  `TimeEndPad(k)` (`transform.py:140-154`)
->  `TimeOut()` on the training split only (`transform.py:175-185`).  Here the three are arguments of ONE kernel,
`ecgvit_patchify_transform`, which reads the raw fp32 records and writes the patch matrix the embedding GEMM consumes;
the normalised / padded / masked signal is never materialised.  Only the random span of `TimeOut` is drawn on the
host, with exactly the reference's torch calls, so a seeded run zeroes the same samples (masking is bit-exact).
"""
import torch

# `config('datasets.PTB-XL.train-stats.denoised')` is what `get_ptbxl_dataset(std_norm=True)` passes (ptb_dataset.py:143);
# the numbers live in the reference's util/config.json:896-957 and are data, not code -- pass them in.


class InputPipeline:
    def __init__(self, normalize=None, pad=None, timeout=None):
        """normalize: dict(mean=[C floats], std=[C floats]) or None        (transform.Normalize)
        pad:       int k -> zero-pad the time axis at the end to a multiple of k, or None/False  (transform.TimeEndPad)
        timeout:   True / (lo, hi) -> zero a random span of every lead while training, or None  (transform.TimeOut)"""
        self.mean = self.std = None
        if normalize is not None:
            mean, std = normalize['mean'], normalize['std']
            assert len(mean) == len(std)  # the reference asserts == 12 (transform.py:26); any lead count works here
            self.mean = torch.tensor(list(mean), dtype=torch.float32)  # np.asarray(..).astype(np.float32)
            self.std = torch.tensor(list(std), dtype=torch.float32)
        if pad:
            assert isinstance(pad, int) and not isinstance(pad, bool), f'If pad, an integer must be provided, got {pad}'
        self.pad = pad or None
        self.timeout = (0.0, 0.5) if timeout is True else (tuple(timeout) if timeout else None)
        self._sampler = None
        if self.timeout is not None:
            self._sampler = torch.distributions.Uniform(low=self.timeout[0], high=self.timeout[1])
        self._dev = {}

    def padded_length(self, length):
        """TimeEndPad pads a FULL extra block when `length` is already a multiple of k (n_pad = k - L % k,
        transform.py:149-151): 2500 -> 2550 at k = 50."""
        if self.pad is None:
            return length
        return length + (self.pad - length % self.pad)

    def draw_spans(self, batch, padded_length):
        """[batch, 2] int32 (start, length) on the host, one TimeOut draw per record in batch order: the same two torch
        RNG calls per record as `TimeOut.__call__` (transform.py:181-183), so `torch.manual_seed(s)` reproduces the
        reference's masks."""
        assert self._sampler is not None
        out = torch.empty(batch, 2, dtype=torch.int32)
        for b in range(batch):
            r = self._sampler.sample().item()
            l_crop = round(r * padded_length)
            idx_strt = torch.randint(high=padded_length - l_crop, size=(1,)).item()
            out[b, 0], out[b, 1] = idx_strt, l_crop
        return out

    def device_stats(self, device):
        """(mean, std) fp32 device tensors, or (None, None)"""
        if self.mean is None:
            return None, None
        key = str(device)
        if key not in self._dev:
            self._dev[key] = (self.mean.to(device), self.std.to(device))
        return self._dev[key]

    def __repr__(self):
        return (f'<{self.__class__.__qualname__} normalize={self.mean is not None} pad={self.pad} '
                f'timeout={self.timeout}>')
