"""Default hot-path options of the reference as an ``argparse.Namespace`` (the reference assembles them from
every layer's ``modify_commandline_options``: options/__init__.py:30-51, swapping_autoencoder_model.py:12-23,
encoder.py:34-37, generator.py:95-102, discriminator.py:8, patch_discriminator.py:14-19,
swapping_autoencoder_optimizer.py:14-21).  bench.py and the tests build their configuration from here."""
from argparse import Namespace


def default_options(**overrides):
    opt = Namespace(
        # experiment / runtime
        name="sae_b200", checkpoints_dir="./checkpoints", isTrain=True, continue_train=False, pretrained_name=None,
        resume_iter="latest", num_gpus=1, batch_size=16, crop_size=256, load_size=256,
        # network selection
        model="swapping_autoencoder", optimizer="swapping_autoencoder",
        netE="StyleGAN2Resnet", netG="StyleGAN2Resnet", netD="StyleGAN2", netPatchD="StyleGAN2",
        use_antialias=True, num_classes=0,
        # loss graph
        spatial_code_ch=8, global_code_ch=2048, lambda_R1=10.0, lambda_patch_R1=1.0, lambda_L1=1.0, lambda_GAN=1.0,
        lambda_PatchGAN=1.0, patch_min_scale=1 / 8, patch_max_scale=1 / 4, patch_num_crops=8,
        patch_use_aggregation=True,
        # encoder
        netE_scale_capacity=1.0, netE_num_downsampling_sp=4, netE_num_downsampling_gl=2, netE_nc_steepness=2.0,
        # generator
        netG_scale_capacity=1.0, netG_num_base_resnet_layers=2, netG_use_noise=True, netG_resnet_ch=256,
        # discriminators
        netD_scale_capacity=1.0, netPatchD_scale_capacity=4.0, netPatchD_max_nc=256 + 128, patch_size=128,
        max_num_tiles=8, patch_random_transformation=False,
        # optimisation
        lr=0.002, beta1=0.0, beta2=0.99, R1_once_every=16,
        # extension (not a reference option): replay each half-step as a CUDA graph (graphs.py)
        cuda_graphs=False,
        # extension: run D (and Dpatch in the discriminator step) once over the concatenated real / rec / mix batch — per-sample
        # identical losses and gradients (tests/test_host_logic.py, 1e-10), same random draws in the same order, a third of the
        # discriminator launches and fuller tiles on the small late layers (+2.9 % images/s measured).  False: three passes,
        # literally as reference models/swapping_autoencoder_model.py:62-98 writes them.
        batch_discriminator_passes=True,
        # extension: losses are read back with one asynchronous copy per half-step and the host only waits when a value is
        # looked at (util.LazyLosses); False: the reference's blocking to_numpy
        async_loss_readback=True,
    )
    for k, v in overrides.items():
        setattr(opt, k, v)
    return opt
