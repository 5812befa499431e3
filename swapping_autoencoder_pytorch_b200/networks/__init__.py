"""Network factory with the reference's name-based lookup (reference models/networks/__init__.py:6-46):
``create_network(opt, "StyleGAN2Resnet", "encoder")`` finds class ``StyleGAN2ResnetEncoder`` (case-insensitive)
in ``networks/encoder.py``."""
import importlib

from .base_network import BaseNetwork

_MODES = ("encoder", "generator", "discriminator", "patch_discriminator")


def find_network_using_name(target_network_name, filename):
    module = importlib.import_module(__name__ + "." + filename)
    wanted = (target_network_name + filename).replace("_", "").lower()
    for name, cls in vars(module).items():
        if isinstance(cls, type) and name.replace("_", "").lower() == wanted and issubclass(cls, BaseNetwork):
            return cls
    raise ValueError("no BaseNetwork subclass %s%s in %s" % (target_network_name, filename, module.__name__))


def modify_commandline_options(parser, is_train):
    opt, _ = parser.parse_known_args()
    for attr, mode in (("netE", "encoder"), ("netG", "generator"), ("netD", "discriminator"),
                       ("netPatchD", "patch_discriminator")):
        name = getattr(opt, attr, None)
        if name is not None:
            parser = find_network_using_name(name, mode).modify_commandline_options(parser, is_train)
    return parser


def create_network(opt, network_name, mode, verbose=False):
    if network_name is None:
        return None
    net = find_network_using_name(network_name, mode)(opt)
    if verbose:
        net.print_architecture(verbose=True)
    return net
