"""The C-ABI library loads on a machine without a GPU and exports every symbol include/osmosis_b200.h declares."""
import ctypes
import os
import re

from osmosis_diffusion_code_b200 import lib as L_

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "osmosis_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(osm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 20
    lib = ctypes.CDLL(L_.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(names) == set(L_.SIGNATURES), set(names) ^ set(L_.SIGNATURES)


def test_abi_version_and_error_string():
    lib = L_.load()
    assert lib.osm_abi_version() == 3
    assert lib.osm_unet_create(None, None) != 0
    assert b"null" in lib.osm_last_error_string()


def test_engine_topology_matches_reference_state_dict_without_gpu():
    import json
    from osmosis_diffusion_code_b200.guided_diffusion.unet import create_model
    from tests.golden.cases import SMALL_UNET
    m = create_model(**SMALL_UNET)
    want = [(k, tuple(s)) for k, s in json.load(open(os.path.join(ROOT, "tests", "golden", "small_unet_param_specs.json")))]
    assert m.param_specs() == want
    assert m.workspace_bytes(2, 32, 32) > 0
