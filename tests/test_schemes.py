"""Scheme presets (host logic, CPU): the field values below were dumped from the unmodified reference's
`auto_round.schemes.PRESET_SCHEMES` (v0.15.0) in the build container."""
import pytest

from auto_round_b200.schemes import PRESETS, parse_scheme

REFERENCE = {
    "W4A16": {"bits": 4, "group_size": 128, "sym": True, "data_type": "int", "act_bits": 16},
    "W2A16": {"bits": 2, "group_size": 128, "sym": True, "data_type": "int", "act_bits": 16},
    "W3A16": {"bits": 3, "group_size": 128, "sym": True, "data_type": "int", "act_bits": 16},
    "W8A16": {"bits": 8, "group_size": 128, "sym": True, "data_type": "int", "act_bits": 16},
    "MXFP4": {"bits": 4, "group_size": 32, "sym": True, "data_type": "mx_fp", "act_bits": 4, "act_group_size": 32,
              "act_sym": True, "act_data_type": "mx_fp", "act_dynamic": True},
    "NVFP4": {"bits": 4, "group_size": 16, "sym": True, "data_type": "nv_fp", "act_bits": 4, "act_group_size": 16,
              "act_sym": True, "act_data_type": "nv_fp4_with_static_gs", "act_dynamic": True},
}


@pytest.mark.parametrize("name", list(REFERENCE))
def test_preset_fields_equal_reference(name):
    ours = dict(PRESETS[name])
    ours.setdefault("sym", True)                      # the FP4 presets are symmetric by construction
    for k, v in REFERENCE[name].items():
        assert ours.get(k) == v, (name, k)
    assert set(ours) <= set(REFERENCE[name]), set(ours) - set(REFERENCE[name])


def test_overrides_and_weight_only_gate():
    s = parse_scheme("W2A16", {"group_size": 32, "sym": False})
    assert (s.bits, s.group_size, s.sym, s.qdq_name) == (2, 32, False, "int_asym")
    assert parse_scheme("NVFP4", {"act_bits": 16, "act_data_type": "float"}).qdq_name == "nv_fp4"
    assert parse_scheme("mxfp4", {"act_bits": 16}).qdq_name == "mx_fp4"
    with pytest.raises(NotImplementedError):          # activation quantisation is outside the hot path: rejected loudly
        parse_scheme("MXFP4")
    with pytest.raises(ValueError):
        parse_scheme("W5A16")
