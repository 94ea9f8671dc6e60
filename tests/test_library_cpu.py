"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol include/nvo_b200.h declares, and the
host-side mirror of the reference interface validates its inputs.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest
import torch


@pytest.fixture(scope="module")
def nv():
    import __graft_entry__ as ge

    ge.build()
    import nerf_vo_b200 as nv

    return nv


def test_library_exports_every_declared_symbol(nv):
    lib = nv._lib.load()
    header = open(nv._lib.HEADER).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(nvo_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 30
    assert declared == nv._lib.declared_symbols()
    for name in declared:
        assert getattr(lib, name) is not None, name
    assert lib.nvo_version() >= 100
    assert lib.nvo_batch_size_granularity() == 128


def test_no_torch_types_in_the_abi(nv):
    code = re.sub(r"/\*.*?\*/", "", open(nv._lib.HEADER).read(), flags=re.S)  # declarations only, comments stripped
    assert "at::" not in code and "torch" not in code.lower() and "tensor" not in code.lower() and "#include <stdint.h>" in code
    out = os.popen(f"nm -D --defined-only {nv._lib.LIB_PATH}").read()
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert all(s in exported for s in nv._lib.declared_symbols())


def test_error_convention_without_gpu(nv):
    """Bad arguments fail with rc != 0 and a message (bindings.cpp:54-55 convention) before any CUDA call is made."""
    lib = nv._lib.load()
    d = nv._lib.make_grid_desc(16, 19, [16.0] * 16)
    d.n_levels = 99
    rc = lib.nvo_grid_forward(ctypes.addressof(d), None, 4, None, None, None)
    assert rc != 0 and b"n_levels" in lib.nvo_last_error()
    m = nv._lib.make_mlp_desc(32, [64, 16], ["relu", "none"])
    assert lib.nvo_mlp_n_params(ctypes.addressof(m)) == 32 * 64 + 64 + 64 * 16 + 16
    assert lib.nvo_mlp_saved_per_sample(ctypes.addressof(m)) == 64
    m.dims[0] = 4096
    assert lib.nvo_mlp_n_params(ctypes.addressof(m)) == -1 and b"width" in lib.nvo_last_error()
    # empty batches are a no-op success on any box
    d = nv._lib.make_grid_desc(16, 19, [16.0] * 16)
    assert lib.nvo_grid_forward(ctypes.addressof(d), None, 0, None, None, None) == 0


def test_host_side_validation(nv):
    spec = nv.ops.GridSpec(16, 12, tuple([16.0] * 16))
    with pytest.raises(RuntimeError, match="CUDA"):
        nv.ops.grid_forward(torch.rand(4, 3), torch.rand(16 << 12, 2), spec)
    with pytest.raises(RuntimeError):
        nv.tcnn_api.Encoding(3, {"otype": "HashGrid", "n_features_per_level": 4})
    with pytest.raises(RuntimeError):
        nv.tcnn_api.Network(32, 16, {"otype": "FullyFusedMLP", "activation": "Swish", "output_activation": "None", "n_neurons": 64, "n_hidden_layers": 1})
    with pytest.raises(NotImplementedError):
        nv.PDFSampler(include_original=True, single_jitter=True)


def test_module_surface_matches_reference_names(nv):
    """Same module / attribute names and state-dict keys as the reference's torch implementation."""
    m = nv.ExtendedNerfactoModel(nv.NerfactoModelConfig(log2_hashmap_size=10), num_train_data=4)
    keys = set(m.state_dict().keys())
    for k in ("field.embedding_appearance.embedding.weight", "field.mlp_base.model.0.hash_table", "field.mlp_base.model.1.layers.1.bias",
              "field.mlp_pred_normals.layers.2.weight", "field.field_head_pred_normals.net.weight", "field.mlp_head.layers.0.weight",
              "proposal_networks.0.encoding.hash_table", "proposal_networks.1.mlp_base.1.layers.1.weight"):
        assert k in keys, k
    assert set(m.get_param_groups()) == {"proposal_networks", "fields"}
    t = nv.tcnn_api.NetworkWithInputEncoding(3, 16, {"otype": "HashGrid", "n_levels": 2, "log2_hashmap_size": 4, "base_resolution": 4, "per_level_scale": 2.0},
                                             {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 16, "n_hidden_layers": 1})
    assert [n for n, _ in t.named_parameters()] == ["params"] and t.n_input_dims == 3 and t.n_output_dims == 16 and t.loss_scale == 1.0


def test_level_scalings_match_oracle(nv):
    import nerfacto_oracle as O

    for gc in (O.GridCfg(), O.GridCfg(5, 16, 128, 17), O.GridCfg(5, 16, 256, 17)):
        assert torch.equal(nv.ops.torch_level_scalings(gc.num_levels, gc.min_res, gc.max_res), O.level_scalings(gc))


def test_new_entry_points_bind_and_validate_without_gpu(nv):
    """The entry points added for the saved Jacobian, the split proposal backward and the grouped exchange: argument lists as the header
    declares them (ctypes argtypes come from the header text, so a drifted call raises TypeError), empty batches are a no-op success and bad
    arguments are rejected with a message before any CUDA call."""
    lib = nv._lib.load()
    A = ctypes.addressof
    err = lambda: lib.nvo_last_error().decode()
    g_tmh = nv._lib.make_grid_desc(16, 19, [16.0] * 16, torch.float32, "tmh")
    g_f32 = nv._lib.make_grid_desc(16, 19, [16.0] * 16)
    g_tmf = nv._lib.make_grid_desc(16, 19, [16.0] * 16, torch.float32, "tmf")
    assert lib.nvo_grid_forward_jac(A(g_tmh), None, 0, None, None, None, None) == 0  # n = 0
    assert lib.nvo_grid_forward_jac(A(g_f32), None, 4, None, None, None, None) != 0 and "NVO_F16_TMH" in err()
    assert lib.nvo_grid_jac_dx(A(g_tmf), None, 0, None, None, 0.0, 1e-12, None) == 0
    assert lib.nvo_grid_jac_dx(A(g_tmf), None, 4, None, None, 0.0, 1e-12, None) != 0 and "null pointer" in err()
    g5 = nv._lib.make_grid_desc(5, 17, [16.0] * 5)
    assert lib.nvo_prop_density_backward_split(A(g5), 16, 0, None, 4, 8, None, None, None, None, None, None) != 0 and "null output" in err()
    assert lib.nvo_prop_density_backward_split(A(g5), 64, 0, None, 4, 8, None, None, None, None, 1, 1) != 0 and "hidden width" in err()
    for off, n, phase in ((2, 8, 0), (0, 6, 0), (0, 8, 7)):  # misaligned offset, misaligned size, phase out of range
        assert lib.nvo_adam_exchange_group(None, off, n, phase, 0, 1, None, None, None, None, None, None, 1e-2, 0.9, 0.999, 1e-15, 1.0, 0) != 0
    assert lib.nvo_adam_exchange_groups2(None, 0, 8, None, None, None, 8, 0, None, None, None, 0, 1, None, None, None, 1e-2, 0.9, 0.999, 1e-15, 1.0) != 0
    assert "second group is empty" in err()
    with pytest.raises(TypeError):
        lib.nvo_grid_jac_dx(A(g_tmf), None, 0)  # too few arguments for the declared signature
    assert lib.nvo_exchange_flag_words() >= 3 * 40
