"""Architecture dictionaries (the hyper-parameter source the reference reads from
``architecture-*.json``; consumed keys per SURVEY 2 #7).  The values below restate
``architecture-vae-vcc2016.json:2-28`` so tests / bench do not depend on a file copied from the
reference; a user's own JSON file is passed through ``ConvVAE(arch)`` verbatim."""
import copy
import json

_VCC2016_VAE = {
    "mode": "VAE",
    "hwc": [513, 1, 1],
    "z_dim": 128,
    "y_dim": 10,
    "y_emb_dim": 128,
    "encoder": {
        "kernel": [[7, 1]] * 5,
        "stride": [[3, 1]] * 5,
        "output": [16, 32, 64, 128, 256],
        "l2-reg": 1e-6,
    },
    "generator": {
        "hwc": [19, 1, 81],
        "merge_dim": 171,
        "kernel": [[9, 1], [7, 1], [7, 1], [1025, 1]],
        "stride": [[3, 1], [3, 1], [3, 1], [1, 1]],
        "output": [32, 16, 8, 1],
        "l2-reg": 1e-6,
    },
    "training": {
        "datadir": "./dataset/vcc2016/bin/Training Set/*/*.bin",
        "batch_size": 16,
        "epoch": 200,
        "lr": 1e-4,
        "beta1": 0.5,
        "beta2": 0.999,
        "max_iter": 60000,
    },
}


def vcc2016_vae_arch():
    return copy.deepcopy(_VCC2016_VAE)


def vawgan_arch():
    """architecture-vawgan-vcc2016.json: the encoder / generator stacks are the ConvVAE ones (the JSON loads unchanged
    through ConvVAE: 939,162 parameters); its extra keys (discriminator, nIterD, lambda, alpha) are not consumed here."""
    a = copy.deepcopy(_VCC2016_VAE)
    a["mode"] = "VAWGAN"
    a["discriminator"] = {"kernel": [[7, 1], [7, 1], [115, 1]], "stride": [[3, 1], [3, 1], [3, 1]], "output": [16, 32, 64],
                          "l2-reg": 1e-6, "merge_dim": 1024, "feature_layer": 1}
    a["training"].update({"nIterD": 5, "lambda": 10, "alpha": 50.0})
    return a


def vawgan_d_stack_arch():
    """cfg5 (BASELINE.json configs[4]): the VAWGAN discriminator conv stack (architecture-vawgan-vcc2016.json:7-14:
    kernels 7 / 7 / 115, stride 3, 16 / 32 / 64 channels, [1,513] -> [64,19]) run as the ENCODER stack of the path's
    graph (conv + Layernorm + lrelu per layer, util/layers.py:47-66), followed by the ConvVAE heads / generator.
    Only the conv stack is specifiable: the VAWGAN model code is absent from the reference snapshot."""
    a = copy.deepcopy(_VCC2016_VAE)
    a["encoder"] = {"kernel": [[7, 1], [7, 1], [115, 1]], "stride": [[3, 1], [3, 1], [3, 1]], "output": [16, 32, 64], "l2-reg": 1e-6}
    return a


def write_arch_json(path, arch=None):
    with open(path, "w") as f:
        json.dump(arch or _VCC2016_VAE, f, indent=4)
