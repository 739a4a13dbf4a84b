"""quantization_config parity (host logic, CPU): the dicts below were written by the unmodified reference's
`quantize_and_save(format="auto_round")` on the tiny Llama (v0.15.0; W4A16 g32) for the stated AutoRound kwargs."""
from auto_round_b200.export import build_quantization_config
from auto_round_b200.schemes import parse_scheme

BASE = {"autoround_version": "0.15.0", "bits": 4, "block_name_to_quantize": "model.layers", "data_type": "int", "group_size": 32,
        "packing_format": "auto_round:auto_gptq", "quant_method": "auto-round", "static_attention_granularity": "tensor",
        "static_kv_granularity": "tensor", "sym": True}


def test_tuned_run_configs_match_reference():
    sc = parse_scheme("W4A16", {"group_size": 32})
    # AutoRound(iters=2)
    assert build_quantization_config(sc, "model.layers", None, 2, 8, 16, 4, tuning={}) == dict(BASE, iters=2)
    # AutoRound(iters=2, lr=0.01, minmax_lr=0.02, enable_minmax_tuning=False, enable_quanted_input=False, not_use_best_mse=True)
    got = build_quantization_config(sc, "model.layers", None, 2, 8, 16, 4,
                                    tuning=dict(lr=0.01, minmax_lr=0.02, enable_minmax_tuning=False, enable_quanted_input=False,
                                                not_use_best_mse=True))
    assert got == dict(BASE, iters=2, lr=0.01, minmax_lr=0.02, enable_minmax_tuning=False, enable_quanted_input=False)
    # defaults are filtered: iters=200 with lr == 1/iters leaves nothing behind
    assert build_quantization_config(sc, "model.layers", None, 200, 128, 2048, 8, tuning=dict(lr=1 / 200)) == BASE
    # RTN (iters=0): tests/golden/rtn_export_*.pt
    assert build_quantization_config(sc, "model.layers", None, 0) == dict(BASE, enable_quanted_input=False)


def test_lm_head_extra_config_matches_reference():
    """`quant_lm_head=True`, W4A16 g32, iters=2: the quantization_config the reference wrote."""
    from auto_round_b200.export import extra_config_entry

    sc = parse_scheme("W4A16", {"group_size": 32})
    ref_extra = {"lm_head": {"act_bits": 16, "act_data_type": "float", "act_dynamic": True, "act_group_size": 32, "act_sym": True,
                             "bits": 4, "data_type": "int", "group_size": 32, "rotation_config": None, "super_bits": None,
                             "super_group_size": None, "sym": True}}
    got = build_quantization_config(sc, "model.layers", {"lm_head": extra_config_entry(sc)}, 2, 8, 16, 4, tuning={})
    assert got == dict(BASE, iters=2, extra_config=ref_extra)
