"""The C-ABI shared library loads without a GPU and exports exactly what include/mmf_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mmf_b200.h")


TRAIN_HEADER = os.path.join(ROOT, "include", "mmf_b200_train.h")


def _declared_functions(header=HEADER):
    src = open(header).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?[A-Za-z_][A-Za-z0-9_]*\s*\*?\s*(mmf_[a-z0-9_]+)\s*\(", src, flags=re.M)
    return sorted(set(names))


@pytest.fixture(scope="module")
def built_lib():
    import __graft_entry__ as entry
    entry.build()
    from mmf_b200 import _abi
    return _abi


def test_header_declares_the_documented_entry_points():
    names = _declared_functions()
    for must in ("mmf_abi_version", "mmf_last_error", "mmf_model_create", "mmf_model_destroy", "mmf_encoder_forward",
                 "mmf_hybrid_step", "mmf_euler_step", "mmf_generate", "mmf_generate_host"):
        assert must in names


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib.LIB_PATH)
    for name in _declared_functions():
        assert hasattr(lib, name), f"{name} is declared in include/mmf_b200.h but not exported"
    # and the Python binding knows about every one of them
    assert sorted(built_lib.EXPORTS) == _declared_functions()
    # the training-step operators (include/mmf_b200_train.h): exported, bound with the right number of arguments
    from mmf_b200 import _train_abi
    declared = _declared_functions(TRAIN_HEADER)
    assert declared == _train_abi.TRAIN_EXPORTS and len(declared) >= 29
    src = re.sub(r"/\*.*?\*/", "", open(TRAIN_HEADER).read(), flags=re.S)
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/mmf_b200_train.h but not exported"
        args = re.search(name + r"\s*\((.*?)\)\s*;", src, flags=re.S).group(1)
        assert len(_train_abi.SIGNATURES[name]) == len([a for a in args.split(",") if a.strip()]), name


def test_abi_version_and_struct_layout(built_lib):
    L = built_lib.lib()
    assert L.mmf_abi_version() == 2
    # struct sizes must equal the C layout (11 x int32; float,float,int32,float,int32,pad,uint64,uint64)
    assert ctypes.sizeof(built_lib.MmfModelDesc) == 44
    assert ctypes.sizeof(built_lib.MmfStepOptions) == 40
    assert ctypes.sizeof(built_lib.MmfWeightRef) == 8 + 8 + 8 + 32


def test_no_cpu_fallback(built_lib):
    """Without a CUDA device the product path must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mmf_b200.mmf import MultiModalFlowBridge
    from mmf_b200.param_spec import make_config
    from mmf_b200 import synthetic
    cfg = make_config("FusedParticleFormer", num_timesteps=2)
    bridge = MultiModalFlowBridge(cfg)
    batch = synthetic.source_batch(2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        bridge.simulate_dynamics(batch)
    with pytest.raises(RuntimeError):
        built_lib.NativeModel(cfg, synthetic.make_state_dict(cfg), torch.device("cpu"))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "multimodal-flows_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f"{f} reaches into oracle/"
