"""confild_b200 -- B200-native CNF decode path of CoNFiLD behind the reference's module API.

    from confild_b200 import SIRENAutodecoder_film          # drop-in for cnf.nf_networks.SIRENAutodecoder_film
    confild_b200.install()                                   # or: patch the reference's module in place

Only the decode hot path lives here: the host-side mirror of the reference
interface (``nf_networks``, ``inference_function``), frame sharding (``distributed``) and the CUDA
sources + C ABI (``csrc/``, ``include/confild_cnf.h``).
"""
from .nf_networks import (BatchLinear, Sine, SIRENAutodecoder_film, SIRENAutodecoder_film_extra_in,  # noqa: F401
                          canonicalize, first_layer_sine_init, sine_init, PRECISION_NOTES)
from .inference_function import decoder, pass_through_model_batch  # noqa: F401
from .folding import fold_normalizers  # noqa: F401
from .dps import GraphedMeasurementNorm, measurement_norm, sensor_rows  # noqa: F401
from .latent_sampler import DDPMSchedule, LatentUNet, generate_fields, sample_latents  # noqa: F401
from .distributed import FusedGatherDecoder, all_gather_frames, decode_frame_sharded, shard_bounds  # noqa: F401

__all__ = [
    "SIRENAutodecoder_film", "SIRENAutodecoder_film_extra_in", "BatchLinear", "Sine",
    "decoder", "pass_through_model_batch", "decode_frame_sharded", "all_gather_frames", "shard_bounds",
    "FusedGatherDecoder", "fold_normalizers", "measurement_norm", "GraphedMeasurementNorm", "sensor_rows", "LatentUNet", "DDPMSchedule", "sample_latents", "generate_fields",
    "install",
]


def install(*modules) -> list:
    """Replace ``SIRENAutodecoder_film`` (and ``_extra_in``) inside already-imported reference modules.

    With no arguments, patches every imported module whose name ends in ``nf_networks`` or
    ``measurements`` and that defines the class (the reference finds it by
    ``getattr(nf_networks, name)`` -- scripts/train.py:230, cnf/inference_function.py:177-180 -- or by
    direct import -- guided_diffusion/measurements.py:7).  Returns the list of patched module names.
    """
    import sys

    if not modules:
        modules = [m for name, m in list(sys.modules.items())
                   if m is not None and name.split(".")[-1] in ("nf_networks", "measurements")
                   and not name.startswith("confild_b200") and hasattr(m, "SIRENAutodecoder_film")]
    patched = []
    for m in modules:
        m.SIRENAutodecoder_film = SIRENAutodecoder_film
        if hasattr(m, "SIRENAutodecoder_film_extra_in"):
            m.SIRENAutodecoder_film_extra_in = SIRENAutodecoder_film_extra_in
        patched.append(m.__name__)
    return patched
