"""Training driver of the hot path: alternating discriminator / generator half-steps with lazy R1.

Restates reference ``optimizers/swapping_autoencoder_optimizer.py:7-119`` (same public methods, same Adam
hyper-parameters and R1 schedule) for boxes without the reference checkout.  The data-parallel gradient
exchange is invisible here: ``MultiGPUModelWrapper`` (parallel.py) all-reduces the active parameter group at
the end of every ``backward()``.
"""
import os

import torch

from . import backend, util


class MultiTensorAdam:
    """``torch.optim.Adam`` for one parameter group as ONE kernel launch (``sae_adam_step``) — the subset of
    the torch optimizer interface the reference driver uses (``zero_grad`` / ``step`` / ``state_dict`` / ``load_state_dict`` /
    ``param_groups``; reference optimizers/swapping_autoencoder_optimizer.py:34-42, :76, :94).  Same arithmetic and the same
    per-parameter semantics: a parameter without a gradient is skipped and keeps its own step count.  Moments live in two
    flat fp32 buffers, step counts on the device, so the update is capturable in a CUDA graph.  ``step(grads=...)`` lets the
    data-parallel path hand in views of the flat all-reduce bucket with ``grad_scale = 1 / world`` (no unpack pass)."""

    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8):
        self.params = list(params)
        self.param_groups = [dict(params=self.params, lr=float(lr), betas=(float(betas[0]), float(betas[1])), eps=float(eps),
                                  weight_decay=0, amsgrad=False, maximize=False)]
        sizes, offsets, total = [], [], 0
        for p in self.params:
            sizes.append(p.numel())
            offsets.append(total)
            total += (p.numel() + 3) // 4 * 4          # every segment 16-byte aligned
        self._sizes, self._offsets, self._total = sizes, offsets, total
        self._dev = None
        self._cache = None

    def _state(self):
        if self._dev is None:
            dev = self.params[0].device
            self.exp_avg = torch.zeros(self._total, dtype=torch.float32, device=dev)
            self.exp_avg_sq = torch.zeros(self._total, dtype=torch.float32, device=dev)
            self.steps = torch.zeros(len(self.params), dtype=torch.float32, device=dev)
            self.offsets_t = torch.tensor(self._offsets, dtype=torch.int64, device=dev)
            self.sizes_t = torch.tensor(self._sizes, dtype=torch.int64, device=dev)
            self._cache = backend.PointerTables(len(self.params), dev)
            self._dev = dev
        return self

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.detach_().zero_()

    @torch.no_grad()
    def step(self, grads=None, grad_scale=1.0):
        st = self._state()
        g = self.param_groups[0]
        if grads is None:
            grads = [p.grad for p in self.params]
        grads = [None if t is None else (t if t.is_contiguous() else t.contiguous()) for t in grads]
        if all(t is None for t in grads):
            return
        params = self.params
        if st.exp_avg.dtype != params[0].dtype:           # fp64 runs of the CPU test-suite
            st.exp_avg, st.exp_avg_sq = st.exp_avg.to(params[0].dtype), st.exp_avg_sq.to(params[0].dtype)
        backend.kernels().adam_step(params, grads, st.offsets_t, st.sizes_t, st.exp_avg, st.exp_avg_sq, st.steps, g["lr"],
                                    g["betas"][0], g["betas"][1], g["eps"], float(grad_scale), self._cache)

    # torch.optim.Adam's on-disk format, so optimizer checkpoints interoperate with the stock optimizer
    def state_dict(self):
        st = self._state()
        state = {}
        for i, (o, n, p) in enumerate(zip(self._offsets, self._sizes, self.params)):
            if float(st.steps[i]) == 0.0:
                continue
            state[i] = {"step": st.steps[i].detach().clone().cpu(), "exp_avg": st.exp_avg[o:o + n].view_as(p).clone(),
                        "exp_avg_sq": st.exp_avg_sq[o:o + n].view_as(p).clone()}
        group = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        group["params"] = list(range(len(self.params)))
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        st = self._state()
        group = sd["param_groups"][0]
        assert len(group["params"]) == len(self.params), "optimizer state belongs to a different parameter list"
        for k in ("lr", "betas", "eps"):
            if k in group:
                self.param_groups[0][k] = tuple(group[k]) if k == "betas" else group[k]
        st.exp_avg.zero_()
        st.exp_avg_sq.zero_()
        st.steps.zero_()
        for i, entry in sd["state"].items():
            i = int(i)
            o, n = self._offsets[i], self._sizes[i]
            st.exp_avg[o:o + n].copy_(entry["exp_avg"].reshape(-1))
            st.exp_avg_sq[o:o + n].copy_(entry["exp_avg_sq"].reshape(-1))
            st.steps[i] = float(entry["step"])


class SwappingAutoencoderOptimizer:
    @staticmethod
    def modify_commandline_options(parser, is_train):
        parser.add_argument("--lr", default=0.002, type=float)
        parser.add_argument("--beta1", default=0.0, type=float)
        parser.add_argument("--beta2", default=0.99, type=float)
        parser.add_argument("--R1_once_every", default=16, type=int,
                            help="lazy R1 regularization: the R1 loss is computed once every this many D iterations")
        return parser

    def __init__(self, model):
        self.opt = opt = model.opt
        self.model = model
        self.train_mode_counter = 0
        self.discriminator_iter_counter = 0
        self.Gparams = model.get_parameters_for_mode("generator")
        self.Dparams = model.get_parameters_for_mode("discriminator")
        on_cuda = bool(self.Gparams and self.Gparams[0].is_cuda)
        # CUDA-graph execution of the half-steps (extension, ``opt.cuda_graphs``): see graphs.py
        self.graphs = None
        if on_cuda and getattr(opt, "cuda_graphs", False):
            from .graphs import HalfStepGraphs
            self.graphs = HalfStepGraphs(self)
        # data parallel: this driver exchanges the gradients itself and hands the flat bucket to Adam (no unpack pass, the
        # 1/world factor folded into the update); the wrapper's end-of-backward averaging into p.grad — what a stock optimizer
        # needs — is switched off
        self.world = getattr(model, "world", 1)
        if self.world > 1:
            model.defer_to_optimizer = True
        self.optimizer_G = MultiTensorAdam(self.Gparams, lr=opt.lr, betas=(opt.beta1, opt.beta2))
        # lazy regularisation correction of lr and betas (StyleGAN2 appendix B; reference :38-42)
        c = opt.R1_once_every / (1 + opt.R1_once_every)
        self.optimizer_D = MultiTensorAdam(self.Dparams, lr=opt.lr * c, betas=(opt.beta1 ** c, opt.beta2 ** c))

    def exchange_and_step(self, optimizer, params):
        """optimizer step of one half-step; with more than one rank: pack -> all-reduce (SUM) -> Adam reading the bucket"""
        if self.world > 1:
            optimizer.step(grads=self.model.reduce_to_bucket(params), grad_scale=1.0 / self.world)
        else:
            optimizer.step()

    @staticmethod
    def set_requires_grad(params, requires_grad):
        for p in params:
            p.requires_grad_(requires_grad)

    def prepare_images(self, data_i):
        return data_i["real_A"]

    def toggle_training_mode(self):
        modes = ["discriminator", "generator"]
        self.train_mode_counter = (self.train_mode_counter + 1) % len(modes)
        return modes[self.train_mode_counter]

    def train_one_step(self, data_i, total_steps_so_far=0):
        """One half-step.  The toggle returns "generator" first, which selects the *discriminator* update
        (reference :59-65) — strict D, G, D, G alternation starting with D."""
        images = self.prepare_images(data_i)
        if self.toggle_training_mode() == "generator":
            losses = self.train_discriminator_one_step(images)
        else:
            losses = self.train_generator_one_step(images)
        return util.to_numpy(losses, lazy=getattr(self.opt, "async_loss_readback", True))

    # ------------------------------------------------------------------ half-step bodies (eager or captured)
    def _generator_body(self, images, step=True):
        self.set_requires_grad(self.Dparams, False)
        self.set_requires_grad(self.Gparams, True)
        self.optimizer_G.zero_grad()
        g_losses, g_metrics = self.model(images, None, None, command="compute_generator_losses")
        sum(v.mean() for v in g_losses.values()).backward()
        if step:
            self.exchange_and_step(self.optimizer_G, self.Gparams)
        g_losses.update(g_metrics)
        return g_losses

    def _discriminator_body(self, images, step=True):
        self.set_requires_grad(self.Dparams, True)
        self.set_requires_grad(self.Gparams, False)
        self.optimizer_D.zero_grad()
        d_losses, d_metrics, sp, gl = self.model(images, command="compute_discriminator_losses")
        sum(v.mean() for v in d_losses.values()).backward()
        # extra outputs travel under "_"-prefixed keys so that eager and captured execution share one flat dict
        d_losses["_sp"], d_losses["_gl"] = sp.detach(), gl.detach()
        d_losses.update({"_metric:" + k: v for k, v in d_metrics.items()})
        if step:
            self.exchange_and_step(self.optimizer_D, self.Dparams)
        return d_losses

    def _r1_body(self, images, step=True):
        self.set_requires_grad(self.Dparams, True)
        self.set_requires_grad(self.Gparams, False)
        self.optimizer_D.zero_grad()
        r1_losses = self.model(images, command="compute_R1_loss")
        (sum(v.mean() for v in r1_losses.values()) * self.opt.R1_once_every).backward()
        if step:
            self.exchange_and_step(self.optimizer_D, self.Dparams)
        return r1_losses

    def _run(self, kind, images):
        """Eager execution of one body, or — with ``opt.cuda_graphs`` on a CUDA device — replay of its CUDA graph."""
        body = {"G": self._generator_body, "D": self._discriminator_body, "R1": self._r1_body}[kind]
        if self.graphs is not None and images.is_cuda:
            return self.graphs.run(kind, body, images)
        return body(images)

    def train_generator_one_step(self, images):
        return self._run("G", images)

    def train_discriminator_one_step(self, images):
        opt = self.opt
        if opt.lambda_GAN == 0.0 and opt.lambda_PatchGAN == 0.0:
            return {}
        self.discriminator_iter_counter += 1
        d_losses = dict(self._run("D", images))
        self.previous_sp, self.previous_gl = d_losses.pop("_sp"), d_losses.pop("_gl")
        d_metrics = {k[len("_metric:"):]: d_losses.pop(k) for k in list(d_losses) if k.startswith("_metric:")}

        needs_r1 = (opt.lambda_R1 > 0.0 or opt.lambda_patch_R1 > 0.0) and \
            self.discriminator_iter_counter % opt.R1_once_every == 0
        if needs_r1:
            d_losses.update(self._run("R1", images))

        d_losses["D_total"] = sum(v.mean() for v in d_losses.values())
        d_losses.update(d_metrics)
        return d_losses

    def get_visuals_for_snapshot(self, data_i):
        with torch.no_grad():
            return self.model(self.prepare_images(data_i), command="get_visuals_for_snapshot")

    # ------------------------------------------------------------------ checkpointing
    def state_dict(self):
        """Adam state of both groups (torch.optim.Adam's format) + the schedule counters.  The reference never saves this
        (optimizers/base_optimizer.py has no state I/O): resuming there restarts Adam's moments from zero."""
        return {"optimizer_G": self.optimizer_G.state_dict(), "optimizer_D": self.optimizer_D.state_dict(),
                "train_mode_counter": self.train_mode_counter, "discriminator_iter_counter": self.discriminator_iter_counter}

    def load_state_dict(self, sd):
        self.optimizer_G.load_state_dict(sd["optimizer_G"])
        self.optimizer_D.load_state_dict(sd["optimizer_D"])
        self.train_mode_counter = int(sd.get("train_mode_counter", 0))
        self.discriminator_iter_counter = int(sd.get("discriminator_iter_counter", 0))

    def _optimizer_path(self, total_steps_so_far=None):
        inner = getattr(self.model, "singlegpu_model", self.model)
        name = "latest_optimizer.pth" if total_steps_so_far is None else "%dk_optimizer.pth" % (total_steps_so_far // 1000)
        return os.path.join(inner._checkpoint_dir(), name)

    def save(self, total_steps_so_far):
        """model checkpoint in the reference's format (models/base_model.py:33-41) + ``<N>k_optimizer.pth`` beside it"""
        self.model.save(total_steps_so_far)
        if getattr(self.model, "rank", 0) == 0:
            path = self._optimizer_path(total_steps_so_far)
            torch.save(self.state_dict(), path)
            link = self._optimizer_path(None)
            if os.path.lexists(link):
                os.remove(link)
            os.symlink(os.path.basename(path), link)

    def load(self, path=None):
        """restore the optimizer state written by ``save`` (missing file: keep the fresh state, like the reference)"""
        path = path or self._optimizer_path(None)
        if not os.path.exists(path):
            return False
        self.load_state_dict(torch.load(path, map_location=str(self.Gparams[0].device)))
        return True
