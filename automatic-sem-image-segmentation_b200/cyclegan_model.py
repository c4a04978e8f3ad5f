"""CycleGanModel on libsemb200: four networks, twelve weight-sharing towers, one train step.

Mirrors /root/reference/Releases/Version 1.2.0/CycleGAN.py: CycleGanModel :512-605, train_step_torch :615-710,
ImagePool :908-964, generator_loss_fn / discriminator_loss_fn :301-308.

train_step_torch semantics kept: `total_gen_loss_a.backward(); total_gen_loss_b.backward()` accumulate the gradients of
BOTH losses into BOTH generators, i.e. one backward of (loss_a + loss_b); the discriminator weight gradients produced
during the generator phase are discarded by `disc.zero_grad()`; the image pool only looks at the first `batch_size`
images it was CONSTRUCTED with.
"""
from __future__ import annotations

import ctypes as C
import random
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib as L
from .engine import Engine
from .gan_nets import DiscriminatorBuilder, GeneratorBuilder

METRICS = ["d_a", "d_b", "d_fake_a", "d_fake_b", "d_real_a", "d_real_b", "g_a", "g_b", "g_adv_a", "g_adv_b", "g_cyc_a",
           "g_cyc_b", "g_id_a", "g_id_b"]


class _Net:
    """One network (flat parameters, Adam state) plus its towers (engines that share those parameters)."""

    def __init__(self, kind: str, h: int, w: int, filters: int, dtype: str, n_res: int, use_tc: bool, seed: int = 0,
                 options: Optional[dict] = None):
        self.kind, self.h, self.w, self.filters, self.dtype, self.n_res, self.use_tc = kind, h, w, filters, dtype, n_res, use_tc
        self.seed = seed
        self.options = dict(options or {})
        self.root: Optional[Engine] = None
        self.towers: Dict[str, tuple] = {}

    def tower(self, name: str, n: int, in_buf=None):
        eng = Engine(n, self.dtype, use_tc=self.use_tc, share=self.root)
        if self.kind == "gen":
            b = GeneratorBuilder(eng, self.h, self.w, self.filters, n_res=self.n_res, in_buf=in_buf, **self.options)
        else:
            b = DiscriminatorBuilder(eng, self.h, self.w, self.filters, in_buf=in_buf, **self.options)
        eng.finalize()
        if self.root is None:
            self.root = eng
            self.names = list(b.creation_names)
            # GlorotUniform kernels, gamma = ones, beta / bias = zeros (CycleGAN.py:125-128).  The reference draws all four
            # networks from ONE stateful SeedGenerator(0) in construction order; here every network has its own seed.
            eng.init_params(self.seed)
        self.towers[name] = (eng, b)
        return eng, b

    # Keras weight protocol (creation order == layer order for these sequential graphs)
    def get_weights(self) -> List[np.ndarray]:
        return [self.root.get_param(n) for n in self.names]

    def set_weights(self, ws):
        if len(ws) != len(self.names):
            raise ValueError(f"expected {len(self.names)} weight arrays, got {len(ws)}")
        for n, w in zip(self.names, ws):
            self.root.set_param(n, np.asarray(w))

    def set_named(self, d):
        for n in self.names:
            self.root.set_param(n, np.asarray(d[n]))


class ImagePool:
    """CycleGAN.py:908-964 on device tensors (storage dtype)."""

    def __init__(self, batch_size: int, pool_size: int = 50, rng=random):
        self.pool_size, self.batch_size, self.rng = pool_size, batch_size, rng
        self.num_imgs, self.images = 0, []

    def query(self, images: torch.Tensor) -> torch.Tensor:
        if self.pool_size == 0:
            return images
        out = []
        for index in range(self.batch_size):
            if index >= images.shape[0]:
                break
            image = images[index:index + 1].clone()
            if self.num_imgs < self.pool_size:
                self.num_imgs += 1
                self.images.append(image)
                out.append(image)
            elif self.rng.uniform(0, 1) > 0.5:
                rid = self.rng.randint(0, self.pool_size - 1)
                out.append(self.images[rid])
                self.images[rid] = image
            else:
                out.append(image)
        return torch.cat(out, 0)


class GeneratorModel:
    """ONE generator tower for inference (`generator(x, training=False)`, CycleGAN.py:266-277): no discriminators, no
    sharing towers, no gradient or optimizer traffic.  InstanceNorm has no inference mode, every sample is independent."""

    def __init__(self, image_shape, batch_size: int, filters: int = 64, dtype: str = "bf16", n_res: int = 9, use_tc: bool = True,
                 **options):
        h, w = image_shape[0], image_shape[1]
        self.n, self.h, self.w = batch_size, h, w
        self.net = _Net("gen", h, w, filters, dtype, n_res, use_tc, options=options)
        self.eng, self.b = self.net.tower("x", batch_size)
        self.x_dev = torch.zeros((batch_size, h, w, 1), dtype=torch.float32, device=self.eng.device)
        self.out_dev = torch.zeros((batch_size,) + tuple(self.b.out_hw) + (1,), dtype=torch.float32, device=self.eng.device)

    def get_weights(self):
        return self.net.get_weights()

    def set_weights(self, ws):
        self.net.set_weights(ws)

    def set_named(self, d):
        self.net.set_named(d)

    def __call__(self, x, training: bool = False) -> np.ndarray:
        """x: (m <= batch, H, W, 1) float32 in [-1, 1]; returns (m, H', W', 1)."""
        x = np.asarray(x, dtype=np.float32)
        m = x.shape[0]
        e, b = self.eng, self.b
        self.x_dev.zero_()
        self.x_dev[:m].copy_(torch.from_numpy(np.ascontiguousarray(x)))
        L.check(e.lib.semb_cast_in(self.x_dev.data_ptr(), 1, C.byref(b.in_buf.view().t), self.n * self.h * self.w, e.dtype, e.stream))
        e.zero_step(False)
        e.forward(True)
        oh, ow = b.out_hw
        L.check(e.lib.semb_cast_out(C.byref(b.out_buf.view().t), self.out_dev.data_ptr(), 1, self.n * oh * ow, e.dtype, e.stream))
        return self.out_dev[:m].cpu().numpy()

    def predict(self, x, batch_size: Optional[int] = None) -> np.ndarray:
        x = np.asarray(x, dtype=np.float32)
        return np.concatenate([self(x[i:i + self.n]) for i in range(0, x.shape[0], self.n)], 0)


class DiscriminatorModel:
    """ONE PatchGAN tower (`get_discriminator(...)`, CycleGAN.py:425-451) as a callable model: forward only."""

    def __init__(self, image_shape, batch_size: int, filters: int = 128, dtype: str = "bf16", use_tc: bool = True, **options):
        h, w = image_shape[0], image_shape[1]
        self.n, self.h, self.w = batch_size, h, w
        self.net = _Net("disc", h, w, filters, dtype, 0, use_tc, options=options)
        self.eng, self.b = self.net.tower("x", batch_size)
        self.x_dev = torch.zeros((batch_size, h, w, 1), dtype=torch.float32, device=self.eng.device)
        oh, ow = self.b.out_hw
        self.out_dev = torch.zeros((batch_size, oh, ow, 1), dtype=torch.float32, device=self.eng.device)

    def get_weights(self):
        return self.net.get_weights()

    def set_weights(self, ws):
        self.net.set_weights(ws)

    def __call__(self, x, training: bool = True) -> np.ndarray:
        e, b = self.eng, self.b
        self.x_dev.copy_(torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32))))
        L.check(e.lib.semb_cast_in(self.x_dev.data_ptr(), 1, C.byref(b.in_buf.view().t), self.n * self.h * self.w, e.dtype, e.stream))
        for op in getattr(b, "noise_ops", []):
            op.frozen = not training
            if not training:
                op.noise.data.zero_()
        e.zero_step(False)
        e.forward(True)
        oh, ow = b.out_hw
        L.check(e.lib.semb_cast_out(C.byref(b.out_buf.view().t), self.out_dev.data_ptr(), 1, self.n * oh * ow, e.dtype, e.stream))
        return self.out_dev.cpu().numpy()


class CycleGanModel:
    def __init__(self, image_shape=(256, 256, 1), batch_size: int = 8, filters: int = 64, dtype: str = "bf16", n_res: int = 9,
                 lambda_cycle_a: float = 10.0, lambda_cycle_b: float = 10.0, lambda_identity_a: float = 0.5,
                 lambda_identity_b: float = 0.5, image_pool_a: Optional[ImagePool] = None, image_pool_b: Optional[ImagePool] = None,
                 label_smoothing_factor: float = 0.0, use_tc: bool = True, use_skip_connection: bool = False,
                 use_resize_convolution: bool = False, gaussian_noise_value: float = 0.0, n_down: int = 3, n_up: int = 3,
                 n_down_disc: int = 2, seed: int = 0, use_cuda_graph: bool = True):
        h, w = image_shape[0], image_shape[1]
        self.h, self.w, self.n, self.dtype = h, w, batch_size, dtype
        self.lc_a, self.lc_b, self.li_a, self.li_b = lambda_cycle_a, lambda_cycle_b, lambda_identity_a, lambda_identity_b
        self.ls = label_smoothing_factor
        self.pool_a = image_pool_a or ImagePool(batch_size, 50)
        self.pool_b = image_pool_b or ImagePool(batch_size, 50)
        npool = min(self.pool_a.batch_size, batch_size) if self.pool_a.pool_size > 0 else batch_size
        self.npool = npool
        n = batch_size
        gopt = {"use_skip_connection": use_skip_connection, "use_resize_convolution": use_resize_convolution, "n_down": n_down, "n_up": n_up}
        dopt = {"gaussian_noise": gaussian_noise_value, "n_down": n_down_disc}
        mk = lambda kind, f, sd: _Net(kind, h, w, f, dtype, n_res, use_tc, seed=sd, options=gopt if kind == "gen" else dopt)
        self.gen_a, self.gen_b = mk("gen", filters, seed), mk("gen", filters, seed + 1)
        self.disc_a, self.disc_b = mk("disc", 2 * filters, seed + 2), mk("disc", 2 * filters, seed + 3)
        # generator towers; fake_b = gen_a(real_a) feeds gen_b and disc_b, so their inputs ALIAS the producer's output
        _, self.GA_ra = self.gen_a.tower("real_a", n)
        _, self.GB_rb = self.gen_b.tower("real_b", n)
        self.real_a, self.real_b = self.GA_ra.in_buf, self.GB_rb.in_buf
        self.fake_b, self.fake_a = self.GA_ra.out_buf, self.GB_rb.out_buf
        self.fake_a.force_acc = self.fake_b.force_acc = True
        _, self.GB_fb = self.gen_b.tower("fake_b", n, in_buf=self.fake_b)
        _, self.GA_fa = self.gen_a.tower("fake_a", n, in_buf=self.fake_a)
        _, self.GB_ra = self.gen_b.tower("same_a", n, in_buf=self.real_a)
        _, self.GA_rb = self.gen_a.tower("same_b", n, in_buf=self.real_b)
        _, self.DA_fa = self.disc_a.tower("fake", n, in_buf=self.fake_a)
        _, self.DB_fb = self.disc_b.tower("fake", n, in_buf=self.fake_b)
        # In the generator phase the discriminators only pass the adversarial gradient back to the fake images: their weight
        # gradients would be discarded (the reference's autograd computes and drops them; SURVEY 8d counts the minimum).
        self.DA_fa.e.skip_wgrad = self.DB_fb.e.skip_wgrad = True
        _, self.DA_real = self.disc_a.tower("real", n, in_buf=self.real_a)
        _, self.DB_real = self.disc_b.tower("real", n, in_buf=self.real_b)
        _, self.DA_pool = self.disc_a.tower("pool", npool)
        _, self.DB_pool = self.disc_b.tower("pool", npool)
        dev = self.gen_a.root.device
        self.dev = dev
        self.lib = self.gen_a.root.lib
        self.sums = torch.zeros(16, dtype=torch.float32, device=dev)
        self.a_dev = torch.zeros((n, h, w, 1), dtype=torch.float32, device=dev)
        self.b_dev = torch.zeros((n, h, w, 1), dtype=torch.float32, device=dev)
        self.a_pin = torch.zeros((n, h, w, 1), dtype=torch.float32).pin_memory()
        self.b_pin = torch.zeros((n, h, w, 1), dtype=torch.float32).pin_memory()
        self.sums_pin = torch.zeros(16, dtype=torch.float32).pin_memory()
        self.learning_rate = 2e-4
        self.beta_1, self.beta_2, self.epsilon = 0.5, 0.999, 1e-7
        self.world_size, self.process_group = 1, None
        self.use_cuda_graph = use_cuda_graph
        self._warm, self._graphs, self._graph_key = False, None, None
        npix = n * h * w
        dpix = n * self.DA_fa.out_hw[0] * self.DA_fa.out_hw[1]
        self._counts = (npix, dpix, self.npool * self.DA_pool.out_hw[0] * self.DA_pool.out_hw[1])

    # ---- helpers -------------------------------------------------------------------------------------
    @property
    def nets(self):
        return {"gen_a": self.gen_a, "gen_b": self.gen_b, "disc_a": self.disc_a, "disc_b": self.disc_b}

    def compile(self, learning_rate: float = 2e-4, beta_1: float = 0.5):
        self.learning_rate, self.beta_1 = learning_rate, beta_1

    def set_distributed(self, process_group=None):
        import torch.distributed as dist
        self.process_group, self.world_size = process_group, dist.get_world_size(process_group)
        from . import dp
        for net in self.nets.values():
            r = net.root
            for t in (r.params.t, r.state.t, r.adam_m, r.adam_v, r.adam_state):
                dp.broadcast_(t, 0, process_group)
            r._pack_dirty = True        # packed tensor-core weight images / virtual stride-2 kernels follow the new master weights

    def _loss(self, eng: Engine, view_a, view_b, target, kind, npix, gscale, slot, acc=0, with_grad=True):
        L.check(self.lib.semb_loss_l1_l2(C.byref(view_a.t), C.byref(view_b.t) if view_b is not None else None, float(target), kind,
                                         npix, 1, float(gscale), C.byref(view_a.g) if with_grad else None, acc,
                                         self.sums.data_ptr() + 4 * slot, eng.dtype, eng.stream))

    def _stage(self, src: torch.Tensor, buf):
        e = self.gen_a.root
        L.check(self.lib.semb_cast_in(src.data_ptr(), 1, C.byref(buf.view().t), src.shape[0] * self.h * self.w, e.dtype, e.stream))

    def _adam(self, net: _Net):
        e = net.root
        e.fold_virtual_grads()          # gradients of the space-to-depth kernels -> Keras-layout gradients (before the all-reduce)
        if self.world_size > 1:
            import torch.distributed as dist
            dist.all_reduce(e.grads, op=dist.ReduceOp.SUM, group=self.process_group)
        e.adam(self.beta_1, self.beta_2, self.epsilon, 1.0 / self.world_size)

    def generate(self, which: str, x) -> np.ndarray:
        """generator(x, training=False) for x NHWC float32 with the model's batch size / shape."""
        x = torch.as_tensor(np.asarray(x, dtype=np.float32))
        b = self.GA_ra if which in ("gen_a", "a") else self.GB_rb
        self.a_dev.copy_(x)
        self._stage(self.a_dev, b.in_buf)
        b.e.zero_step(False)
        b.e.forward(True)            # InstanceNorm has no inference mode: statistics are always per sample
        out = torch.zeros((self.n, self.h, self.w, 1), dtype=torch.float32, device=self.dev)
        L.check(self.lib.semb_cast_out(C.byref(b.out_buf.view().t), out.data_ptr(), 1, self.n * self.h * self.w, b.e.dtype, b.e.stream))
        return out.cpu().numpy()

    # ---- the step ------------------------------------------------------------------------------------------
    def train_step(self, batch) -> Dict[str, float]:
        real_a, real_b = batch

        def host(arr, pin):
            # a caller-pinned float32 tensor is uploaded as it is (zero copy); anything else is staged through pinned memory
            if isinstance(arr, torch.Tensor) and arr.device.type == "cpu" and arr.dtype == torch.float32 and arr.is_contiguous() \
                    and arr.is_pinned() and tuple(arr.shape) == tuple(pin.shape):
                return arr
            pin.copy_(torch.as_tensor(np.ascontiguousarray(arr, dtype=np.float32)))
            return pin

        self.a_dev.copy_(host(real_a, self.a_pin), non_blocking=True)
        self.b_dev.copy_(host(real_b, self.b_pin), non_blocking=True)
        self.step_device()
        self.sums_pin.copy_(self.sums, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._metrics()

    def step_device(self):
        """One train_step_torch (CycleGAN.py:615-710) with the inputs already in a_dev / b_dev.  The first call runs
        eagerly (it also allocates the lazy gradient buffers); from the second call on each phase is ONE CUDA-graph
        replay -- the step has ~1600 launches and is otherwise bound by the host's launch rate.  The image pools sit
        between the two graphs: their swaps are decided by the host RNG, exactly like the reference's Python pool."""
        for net in self.nets.values():
            net.root.lr.fill_(self.learning_rate)         # outside the graphs: the schedule may change it
        use_graph = self.use_cuda_graph and self.world_size == 1
        key = (self.beta_1, self.beta_2, self.epsilon, self.lc_a, self.lc_b, self.li_a, self.li_b, self.ls)
        if use_graph and self._warm and (self._graphs is None or self._graph_key != key):
            g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                self._gen_phase()
            with torch.cuda.graph(g2, pool=g1.pool()):
                self._disc_phase()
            self._graphs, self._graph_key = (g1, g2), key
        if use_graph and self._graphs is not None:
            self._graphs[0].replay()
            self._query_pools()
            self._graphs[1].replay()
        else:
            self._gen_phase()
            self._query_pools()
            self._disc_phase()
            self._warm = True

    def _gen_phase(self):
        n, h, w = self.n, self.h, self.w
        npix = n * h * w
        G = [self.GA_ra, self.GB_rb, self.GB_fb, self.GA_fa, self.GB_ra, self.GA_rb]
        self.sums.zero_()
        self._stage(self.a_dev, self.real_a)
        self._stage(self.b_dev, self.real_b)
        for net in self.nets.values():
            net.root.zero_step(zero_grads=True)
        # ---------------- generator phase: 6 generator + 2 discriminator forwards
        for b in G + [self.DA_fa, self.DB_fb]:
            if b.e.share is not None:
                b.e.zero_step(False)
            b.e.forward(True)
        one = (1.0 - self.ls) + self.ls / 2
        dpix = n * self.DA_fa.out_hw[0] * self.DA_fa.out_hw[1]
        # adversarial (LSGAN): adv_a uses disc_b(fake_b), adv_b uses disc_a(fake_a)
        self._loss(self.DB_fb.e, self.DB_fb.out_buf.view(), None, one, 1, dpix, 1.0 / dpix, 8)
        self._loss(self.DA_fa.e, self.DA_fa.out_buf.view(), None, one, 1, dpix, 1.0 / dpix, 9)
        # cycle: cycle_loss_a = MAE(real_b, cycled_b)*lambda_a, cycled_b = gen_a(fake_a)
        self._loss(self.GA_fa.e, self.GA_fa.out_buf.view(), self.real_b.view(), 0, 0, npix, self.lc_a / npix, 10)
        self._loss(self.GB_fb.e, self.GB_fb.out_buf.view(), self.real_a.view(), 0, 0, npix, self.lc_b / npix, 11)
        # identity: id_loss_a = MAE(real_b, same_b)*lambda_a*lambda_id_a, same_b = gen_a(real_b)
        self._loss(self.GA_rb.e, self.GA_rb.out_buf.view(), self.real_b.view(), 0, 0, npix, self.lc_a * self.li_a / npix, 12)
        self._loss(self.GB_ra.e, self.GB_ra.out_buf.view(), self.real_a.view(), 0, 0, npix, self.lc_b * self.li_b / npix, 13)
        # fake_a / fake_b receive gradients from several consumers
        self.fake_a.grad_tensor().zero_()
        self.fake_b.grad_tensor().zero_()
        for b in [self.DB_fb, self.DA_fa, self.GB_fb, self.GA_fa, self.GB_ra, self.GA_rb, self.GA_ra, self.GB_rb]:
            b.e.backward()
        self._adam(self.gen_a)
        self._adam(self.gen_b)

    def _query_pools(self):
        # ---------------- image pools (generator outputs are detached; the host RNG decides the swaps)
        pa = self.pool_a.query(self.fake_a.data)
        pb = self.pool_b.query(self.fake_b.data)
        self.DA_pool.in_buf.data.copy_(pa)
        self.DB_pool.in_buf.data.copy_(pb)

    def _disc_phase(self):
        # ---------------- discriminator phase (weights of D unchanged so far)
        n = self.n
        one = (1.0 - self.ls) + self.ls / 2
        zero = self.ls / 2
        npix = n * self.h * self.w
        dpix = n * self.DA_fa.out_hw[0] * self.DA_fa.out_hw[1]
        for net in (self.disc_a, self.disc_b):
            net.root.zero_grads()
        D = [self.DA_real, self.DA_pool, self.DB_real, self.DB_pool]
        for b in D:
            b.e.zero_step(False)
            b.e.forward(True)
        ppix = self.npool * self.DA_pool.out_hw[0] * self.DA_pool.out_hw[1]
        # total = 0.5*(MSE(1, real) + MSE(0, fake)); the 0.5 is folded into the gradient scale
        self._loss(self.DA_real.e, self.DA_real.out_buf.view(), None, one, 1, dpix, 0.5 / dpix, 4)
        self._loss(self.DA_pool.e, self.DA_pool.out_buf.view(), None, zero, 1, ppix, 0.5 / ppix, 2)
        self._loss(self.DB_real.e, self.DB_real.out_buf.view(), None, one, 1, dpix, 0.5 / dpix, 5)
        self._loss(self.DB_pool.e, self.DB_pool.out_buf.view(), None, zero, 1, ppix, 0.5 / ppix, 3)
        for b in D:
            b.e.backward()
        self._adam(self.disc_a)
        self._adam(self.disc_b)
        self._counts = (npix, dpix, ppix)

    def _metrics(self) -> Dict[str, float]:
        s = self.sums_pin
        npix, dpix, ppix = self._counts
        m = {"d_fake_a": float(s[2]) / ppix, "d_fake_b": float(s[3]) / ppix, "d_real_a": float(s[4]) / dpix, "d_real_b": float(s[5]) / dpix,
             "g_adv_a": float(s[8]) / dpix, "g_adv_b": float(s[9]) / dpix,
             "g_cyc_a": float(s[10]) / npix * self.lc_a, "g_cyc_b": float(s[11]) / npix * self.lc_b,
             "g_id_a": float(s[12]) / npix * self.lc_a * self.li_a, "g_id_b": float(s[13]) / npix * self.lc_b * self.li_b}
        m["d_a"] = 0.5 * (m["d_real_a"] + m["d_fake_a"])
        m["d_b"] = 0.5 * (m["d_real_b"] + m["d_fake_b"])
        m["g_a"] = m["g_adv_a"] + m["g_cyc_a"] + m["g_id_a"]
        m["g_b"] = m["g_adv_b"] + m["g_cyc_b"] + m["g_id_b"]
        return {k: m[k] for k in METRICS}
