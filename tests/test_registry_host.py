"""CPU: the reference-facing registries and the YAML loader (no device work): every shipped config builds its sampler,
operator, noiser and conditioning method with the reference's names and string-typed YAML values,
unknown names raise NameError like the reference, and the respaced schedule tables match the oracle's."""
import os

import numpy as np
import pytest

from oracle import osmosis_oracle as orc
from osmosis_diffusion_code_b200.guided_diffusion.gaussian_diffusion import create_sampler, get_sampler
from osmosis_diffusion_code_b200.guided_diffusion.measurements import get_operator, get_noise
from osmosis_diffusion_code_b200.guided_diffusion.condition_methods import get_conditioning_method
from osmosis_diffusion_code_b200.guided_diffusion.posterior_mean_variance import get_mean_processor, get_var_processor, coefficient_table
from osmosis_diffusion_code_b200.osmosis_utils.utils import arguments_from_file

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GUIDED = ["osmosis_sample_config.yaml", "osmosis_simulation_sample_config.yaml", "osmosis_haze_sample_config.yaml"]


@pytest.mark.parametrize("name", GUIDED)
def test_guided_configs_build(name):
    a = arguments_from_file(os.path.join(ROOT, "configs", name))
    s = create_sampler(**a.diffusion)
    assert type(s).__name__ == "DDPM" and s.num_timesteps == int(a.diffusion["timestep_respacing"])
    tab = orc.make_tables(a.diffusion["steps"], a.diffusion["noise_schedule"], a.diffusion["timestep_respacing"])
    assert np.array_equal(s.betas, tab.betas) and list(s.timestep_map) == list(tab.timestep_map)
    ct = coefficient_table(s.betas)
    assert ct.shape == (s.num_timesteps, 12) and ct[5, 6] == np.float32(tab.alphas_cumprod[5]) and ct[0, 7] == 1.0
    assert np.isneginf(ct[0, 8]) and ct[0, 10] == 1.0 and ct[7, 9] == np.float32(np.log(s.betas[7]))
    cond_cls = get_conditioning_method.__globals__["__CONDITIONING_METHOD__"][a.conditioning["method"]]
    assert cond_cls.__name__ == "PosteriorSamplingOsmosis"
    assert get_noise(**a.measurement["noise"]).__name__ == a.measurement["noise"]["name"]


def test_rgb_guidance_config_builds_without_a_device():
    a = arguments_from_file(os.path.join(ROOT, "configs", "rgb_guidance_sample_config.yaml"))
    s = create_sampler(**a.diffusion)
    assert s.mean_processor.clip_denoised and s.ddim_eta is None
    op = get_operator(device="cpu", **a.measurement["operator"])
    cond = get_conditioning_method(a.conditioning["method"], op, get_noise(**a.measurement["noise"]), **a.conditioning["params"])
    assert type(cond).__name__ == "PosteriorSampling" and [round(float(v), 3) for v in cond.scale] == [3.0, 3.0, 3.0, 0.1]
    d = dict(a.diffusion); d["sampler"] = "ddim"; d["timestep_respacing"] = "ddim25"
    s2 = create_sampler(**d)
    assert type(s2).__name__ == "DDIM" and s2.num_timesteps == 25 and s2.ddim_eta == 0.0


def test_unknown_names_raise_like_the_reference():
    for fn, arg in ((get_sampler, "plms"), (get_mean_processor, "velocity"), (get_var_processor, "fixed_tiny")):
        with pytest.raises(NameError):
            fn(arg)
    b = np.linspace(1e-4, 0.02, 10)
    for name in ("epsilon", "start_x", "previous_x"):
        assert get_mean_processor(name, betas=b, dynamic_threshold=False, clip_denoised=True).flags & 1
    assert [get_var_processor(n, betas=b).flags for n in ("learned_range", "learned", "fixed_small", "fixed_large")] == [0, 0x100, 0x200, 0x300]
    with pytest.raises(NameError):
        get_operator("super_resolution", device="cpu")
    with pytest.raises(NameError):
        get_noise("poisson", rate=1.0)
    with pytest.raises(NameError):
        get_conditioning_method("mcg", None, None)
    with pytest.raises(NotImplementedError):
        get_mean_processor("epsilon", betas=np.linspace(1e-4, 0.02, 10), dynamic_threshold=True, clip_denoised=False)
