"""Training driver of the hot path: alternating discriminator / generator half-steps with lazy R1.

Restates reference ``optimizers/swapping_autoencoder_optimizer.py:7-119`` (same public methods, same Adam
hyper-parameters and R1 schedule) for boxes without the reference checkout.  The data-parallel gradient
exchange is invisible here: ``MultiGPUModelWrapper`` (parallel.py) all-reduces the active parameter group at
the end of every ``backward()``.
"""
import torch

from . import util


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
        fused = dict(fused=True, capturable=self.graphs is not None) if on_cuda else {}
        self.optimizer_G = torch.optim.Adam(self.Gparams, lr=opt.lr, betas=(opt.beta1, opt.beta2), **fused)
        # lazy regularisation correction of lr and betas (StyleGAN2 appendix B; reference :38-42)
        c = opt.R1_once_every / (1 + opt.R1_once_every)
        self.optimizer_D = torch.optim.Adam(self.Dparams, lr=opt.lr * c, betas=(opt.beta1 ** c, opt.beta2 ** c), **fused)

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
        return util.to_numpy(losses)

    # ------------------------------------------------------------------ half-step bodies (eager or captured)
    def _generator_body(self, images, step=True):
        self.set_requires_grad(self.Dparams, False)
        self.set_requires_grad(self.Gparams, True)
        self.optimizer_G.zero_grad()
        g_losses, g_metrics = self.model(images, None, None, command="compute_generator_losses")
        sum(v.mean() for v in g_losses.values()).backward()
        if step:
            self.optimizer_G.step()
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
            self.optimizer_D.step()
        return d_losses

    def _r1_body(self, images, step=True):
        self.set_requires_grad(self.Dparams, True)
        self.set_requires_grad(self.Gparams, False)
        self.optimizer_D.zero_grad()
        r1_losses = self.model(images, command="compute_R1_loss")
        (sum(v.mean() for v in r1_losses.values()) * self.opt.R1_once_every).backward()
        if step:
            self.optimizer_D.step()
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

    def save(self, total_steps_so_far):
        self.model.save(total_steps_so_far)
