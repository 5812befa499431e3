"""Helper of tests/test_checkpoint.py (build container only): the REFERENCE's own SwappingAutoencoderModel, on the CPU, writes a
checkpoint with its own BaseModel.save (models/base_model.py:33-41) and the encoder outputs for a seeded image next to it.
usage: python tests/ref_checkpoint_driver.py <repo root> <reference root> <checkpoints dir>"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT, REF, OUT = sys.argv[1:4]
sys.path.insert(0, ROOT)
os.environ["SAE_REFERENCE_ROOT"] = REF

from oracle import ref_import  # noqa: E402
from oracle.fixtures import TINY, perturbed_state_dict, rnd  # noqa: E402
from swapping_autoencoder_pytorch_b200 import default_options  # noqa: E402

R = ref_import.import_reference()
opt = default_options(**dict(TINY, checkpoints_dir=OUT, name="written_by_reference"))
model = R.sae_model.SwappingAutoencoderModel(opt)
with contextlib.redirect_stdout(io.StringIO()):
    model.initialize()
sd = perturbed_state_dict(opt, dtype=torch.float32, param_seed=7, bias_seed=11)
missing, unexpected = model.load_state_dict(sd, strict=False)
assert not unexpected, unexpected
os.makedirs(os.path.join(OUT, opt.name), exist_ok=True)
model.save(7000)                                   # the reference's own save: 7k_checkpoint.pth + latest_checkpoint.pth
real = rnd(900, 2, 3, 64, 64).clamp(-1, 1).float()
with torch.no_grad():
    sp, gl = model.E(real)
np.savez(os.path.join(OUT, "reference_encoder_outputs.npz"), sp=sp.numpy(), gl=gl.numpy())
print("ok", len(model.state_dict()))
