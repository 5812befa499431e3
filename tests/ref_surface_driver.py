"""TEST INFRASTRUCTURE (CPU, build container only).  Runs the REFERENCE'S OWN swapping_autoencoder_model.py,
swapping_autoencoder_optimizer.py and network files (encoder / generator / discriminator / patch_discriminator) over this
repo's operator surface — the sys.modules aliasing of INTEGRATION.md §1 — with the kernel interface emulated on the CPU,
and, on the same parameters and RNG seeds, this repo's restated callers.  Prints one JSON line.  Own process because it
re-binds ``models.networks.stylegan2_*`` in sys.modules."""
import contextlib
import io
import json
import sys
import types

import torch

ROOT, REF = sys.argv[1], sys.argv[2]
sys.path.insert(0, ROOT)
torch.set_default_dtype(torch.float64)

from swapping_autoencoder_pytorch_b200 import backend, default_options      # noqa: E402
import swapping_autoencoder_pytorch_b200 as S                               # noqa: E402
import swapping_autoencoder_pytorch_b200.stylegan2_op as _op                # noqa: E402
import swapping_autoencoder_pytorch_b200.stylegan2_layers as _layers        # noqa: E402
from tests.cpu_emulation import EmulatedKernels                             # noqa: E402
from oracle.fixtures import TINY, perturbed_state_dict, rnd                 # noqa: E402

backend.set_kernels(EmulatedKernels())


def stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


stub("func_timeout", func_timeout=lambda *a, **k: None, FunctionTimedOut=type("FunctionTimedOut", (Exception,), {}))
tags = stub("dominate.tags", **{t: (lambda *a, **k: None) for t in ("meta", "h3", "table", "tr", "td", "p", "a", "img", "br")})
stub("dominate", document=lambda *a, **k: None, tags=tags)
stub("visdom", Visdom=lambda *a, **k: None)
sys.path.insert(0, REF)
# ---- the two aliases, BEFORE the reference's "models" package is imported
sys.modules["models.networks.stylegan2_op"] = _op
sys.modules["models.networks.stylegan2_layers"] = _layers
import models                                                               # noqa: E402  (the reference's package)
import models.swapping_autoencoder_model as ref_model_mod                   # noqa: E402
import optimizers.swapping_autoencoder_optimizer as ref_opt_mod            # noqa: E402
import models.networks.encoder as ref_enc                                   # noqa: E402

assert ref_model_mod.__file__.startswith(REF) and ref_opt_mod.__file__.startswith(REF) and ref_enc.__file__.startswith(REF)
assert ref_enc.ConvLayer is _layers.ConvLayer            # the reference's network files build from THIS repo's classes

opt = default_options(**dict(TINY, R1_once_every=1))
sd = perturbed_state_dict(opt, dtype=torch.float64)
real = rnd(900, 2, 3, 64, 64).clamp(-1, 1)


def run(make_model, make_trainer):
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        wrapper = make_model(opt)
    inner = wrapper.singlegpu_model
    missing, unexpected = inner.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    trainer = make_trainer(wrapper)
    torch.manual_seed(1)
    out = {}
    for tag in ("D", "G"):
        losses = trainer.train_one_step({"real_A": real}, 0)
        out.update({tag + "/" + k: float(v) for k, v in losses.items()})
    out["param_checksum"] = float(sum(p.double().abs().sum() for p in inner.parameters()))
    return out


ref = run(models.create_model, ref_opt_mod.SwappingAutoencoderOptimizer)
ours = run(S.create_model, lambda w: S.create_optimizer(opt, w))
print(json.dumps({"reference_files": ref, "restated_callers": ours}))
