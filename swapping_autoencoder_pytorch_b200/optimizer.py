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
        fused = dict(fused=True) if (self.Gparams and self.Gparams[0].is_cuda) else {}
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

    def train_generator_one_step(self, images):
        self.set_requires_grad(self.Dparams, False)
        self.set_requires_grad(self.Gparams, True)
        self.optimizer_G.zero_grad()
        g_losses, g_metrics = self.model(images, None, None, command="compute_generator_losses")
        sum(v.mean() for v in g_losses.values()).backward()
        self.optimizer_G.step()
        g_losses.update(g_metrics)
        return g_losses

    def train_discriminator_one_step(self, images):
        opt = self.opt
        if opt.lambda_GAN == 0.0 and opt.lambda_PatchGAN == 0.0:
            return {}
        self.set_requires_grad(self.Dparams, True)
        self.set_requires_grad(self.Gparams, False)
        self.discriminator_iter_counter += 1
        self.optimizer_D.zero_grad()
        d_losses, d_metrics, sp, gl = self.model(images, command="compute_discriminator_losses")
        self.previous_sp, self.previous_gl = sp.detach(), gl.detach()
        sum(v.mean() for v in d_losses.values()).backward()
        self.optimizer_D.step()

        needs_r1 = (opt.lambda_R1 > 0.0 or opt.lambda_patch_R1 > 0.0) and \
            self.discriminator_iter_counter % opt.R1_once_every == 0
        if needs_r1:
            self.optimizer_D.zero_grad()
            r1_losses = self.model(images, command="compute_R1_loss")
            d_losses.update(r1_losses)
            (sum(v.mean() for v in r1_losses.values()) * opt.R1_once_every).backward()
            self.optimizer_D.step()

        d_losses["D_total"] = sum(v.mean() for v in d_losses.values())
        d_losses.update(d_metrics)
        return d_losses

    def get_visuals_for_snapshot(self, data_i):
        with torch.no_grad():
            return self.model(self.prepare_images(data_i), command="get_visuals_for_snapshot")

    def save(self, total_steps_so_far):
        self.model.save(total_steps_so_far)
