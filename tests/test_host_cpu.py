"""CPU: host-side logic that needs no GPU -- registries, config mirrors, schema, flat-store ordering, masks."""
import pytest
import torch


def test_registries_and_config_defaults_mirror_reference():
    import liteasr_b200.criterions as C
    import liteasr_b200.models as M
    from liteasr_b200.criterions.hybrid_ctc_attn import HybridCTCLossConfig
    from liteasr_b200.models.u2 import U2Config
    assert "U2" in M.MODEL_REGISTRY and "hybrid_ctc" in C.CRITERION_REGISTRY
    c = U2Config(input_dim=80, vocab_size=100)
    # models/u2.py:35-67 defaults
    assert (c.name, c.dropout_rate, c.use_rel, c.enc_dim, c.enc_ff_dim, c.enc_attn_heads, c.enc_layers, c.activation) == \
        ("U2", 0.0, True, 256, 2048, 4, 12, "swish")
    assert (c.dec_dim, c.dec_ff_dim, c.dec_attn_heads, c.dec_layers) == (256, 2048, 4, 6)
    assert c.enc_attn_dropout_rate == "${model.enc_dropout_rate}"  # II(...) interpolation strings kept
    h = HybridCTCLossConfig()
    assert (h.name, h.padding_idx, h.smoothing, h.normalize_length, h.ctc_weight) == ("hybrid_ctc", -1, 0.0, False, 0.0)


def test_build_model_and_criterion_through_registry():
    from types import SimpleNamespace as NS
    import liteasr_b200.criterions as C
    import liteasr_b200.models as M
    task = NS(feat_dim=80, vocab_size=37)
    m = M.build_model(NS(name="U2", enc_layers=1, dec_layers=1, enc_dim=64, dec_dim=64, enc_attn_heads=1, dec_attn_heads=1,
                         enc_ff_dim=96, dec_ff_dim=96), task)
    assert m.sos == m.eos == 36 and m.blank == 0 and m.ignore == -1
    crit = C.build_criterion(NS(name="hybrid_ctc", ctc_weight=0.3, smoothing=0.1), task)
    assert crit.cfg.vocab_size == 37 and crit.cfg.ctc_weight == 0.3


def test_state_dict_schema_matches_reference_keys():
    from liteasr_b200.models.u2 import U2, U2Config
    from liteasr_b200.schema import U2Dims, u2_schema
    dims = U2Dims(80, 50, 128, 256, 2, 2, 128, 256, 2, 2)
    m = U2(U2Config(**dims.__dict__))
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {n: tuple(s) for n, s, _ in u2_schema(dims)}
    from liteasr_b200.utils.synthetic import synth_state_dict
    m.load_state_dict(synth_state_dict(dims, 1), strict=True)


def test_unsupported_configs_fail_loudly():
    from liteasr_b200.models.u2 import U2, U2Config
    with pytest.raises(ValueError):
        U2(U2Config(input_dim=80, vocab_size=50, enc_layers=1, dec_layers=1, dropout_rate=1.0))
    with pytest.raises(NotImplementedError):
        U2(U2Config(input_dim=80, vocab_size=50, use_rel=False))


def test_reference_yaml_dropout_rates_construct_and_interpolate():
    """config/model/my_U2.yaml: dropout_rate 0.1, every ${model...} rate follows it, the three attention rates are 0.0."""
    from liteasr_b200.dropout import has_dropout
    from liteasr_b200.models.u2 import U2, U2Config
    m = U2(U2Config(input_dim=80, vocab_size=50, enc_layers=1, dec_layers=1, dropout_rate=0.1, enc_attn_dropout_rate=0.0,
                    dec_self_attn_dropout_rate=0.0, dec_src_attn_dropout_rate=0.0))
    e, d = m.encoder.enc_layers[0], m.decoder.dec_layers[0]
    assert (e.dropout_rate, e.feed_forward.dropout_rate, e.feed_forward_macaron.dropout_rate, e.self_attn.dropout_rate) == (0.1, 0.1, 0.1, 0.0)
    assert (d.dropout_rate, d.feed_forward.dropout_rate, d.self_attn.dropout_rate, d.src_attn.dropout_rate) == (0.1, 0.1, 0.0, 0.0)
    assert m.encoder.pe.dropout_rate == 0.1 and m.decoder.pe.dropout_rate == 0.1 and m.ctc.dropout_rate == 0.1
    assert has_dropout(m) and has_dropout(m.ctc)
    m0 = U2(U2Config(input_dim=80, vocab_size=50, enc_layers=1, dec_layers=1))
    assert not has_dropout(m0)


def test_cpu_forward_fails_loudly():
    from liteasr_b200.models.u2 import U2, U2Config
    m = U2(U2Config(input_dim=80, vocab_size=50, enc_layers=1, dec_layers=1, enc_dim=64, dec_dim=64, enc_attn_heads=1,
                    dec_attn_heads=1, enc_ff_dim=96, dec_ff_dim=96))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 40, 80), torch.tensor([40]), torch.tensor([[3]]), torch.tensor([1]))


def test_flat_order_groups_qkv():
    from liteasr_b200.models.u2 import U2, U2Config
    from liteasr_b200.store import _ordered_named_parameters
    m = U2(U2Config(input_dim=80, vocab_size=50, enc_layers=1, dec_layers=1, enc_dim=64, dec_dim=64, enc_attn_heads=1,
                    dec_attn_heads=1, enc_ff_dim=96, dec_ff_dim=96))
    named, tight = _ordered_named_parameters(m)
    names = [n for n, _ in named]
    assert sorted(names) == sorted(n for n, _ in m.named_parameters())
    i = names.index("encoder.enc_layers.0.self_attn.linear_q.weight")
    assert names[i:i + 6] == [f"encoder.enc_layers.0.self_attn.linear_{x}.{k}" for k in ("weight", "bias") for x in "qkv"]
    assert {names[i], names[i + 1], names[i + 3], names[i + 4]} <= tight and names[i + 2] not in tight


def test_qkv_spans_are_contiguous_for_widths_that_are_not_multiples_of_64():
    """the fused q/k/v projection reads the three biases / weights as ONE span; with a 64-element padding
    after every parameter that only held for d % 64 == 0.  The group is now packed tightly for any d % 4 == 0."""
    from liteasr_b200.models.u2 import U2, U2Config
    from liteasr_b200.store import ParamStore
    m = U2(U2Config(input_dim=80, vocab_size=50, enc_layers=1, dec_layers=1, enc_dim=144, dec_dim=144, enc_attn_heads=4,
                    dec_attn_heads=4, enc_ff_dim=96, dec_ff_dim=96))
    want = {n: p.detach().clone() for n, p in m.named_parameters()}
    st = ParamStore(m, torch.device("cpu"), "fp32")
    d = 144
    for pre in ("encoder.enc_layers.0.self_attn", "decoder.dec_layers.0.self_attn"):
        w = st.w(pre + ".linear_q.weight", 3 * d, d)
        b = st.p_span(pre + ".linear_q.bias", 3 * d)
        for j, x in enumerate("qkv"):
            assert torch.equal(w[j * d:(j + 1) * d], want[f"{pre}.linear_{x}.weight"])
            assert torch.equal(b[j * d:(j + 1) * d], want[f"{pre}.linear_{x}.bias"])
            assert st.off[f"{pre}.linear_{x}.weight"][0] % 4 == 0 and st.off[f"{pre}.linear_{x}.bias"][0] % 4 == 0
    pre = "decoder.dec_layers.0.src_attn"
    kv = st.p_span(pre + ".linear_k.bias", 2 * d)
    assert torch.equal(kv[:d], want[pre + ".linear_k.bias"]) and torch.equal(kv[d:], want[pre + ".linear_v.bias"])
    lo, hi = st.range_of("decoder.")
    assert lo % 64 == 0 and hi % 64 == 0 and hi <= st.numel
    for n, p in m.named_parameters():  # parameters are views of the flat buffer and keep their values
        assert torch.equal(p, want[n]) and p.data_ptr() == st.flat.data_ptr() + 4 * st.off[n][0]


def test_masks_and_synthetic_batch_contract():
    from liteasr_b200.utils.mask import padding_mask, triangle_mask
    from liteasr_b200.utils.synthetic import pred_len, synth_batch
    assert padding_mask(torch.tensor([5, 3, 1])).int().tolist() == [[0, 0, 0, 0, 0], [0, 0, 0, 1, 1], [0, 1, 1, 1, 1]]
    assert triangle_mask(3, 5, diagonal=2).int().tolist() == [[0, 0, 1, 1, 1], [0, 0, 0, 1, 1], [0, 0, 0, 0, 1]]
    xs, xlens, ys, ylens = synth_batch(8, 500, 30, 500, seed=42)
    assert xs.shape == (8, 500, 80) and xlens[0] == 500 and (xlens[:-1] >= xlens[1:]).all()
    assert (ys[torch.arange(8), ylens - 1] > 0).all() and ((ys == -1) == (torch.arange(ys.shape[1])[None] >= ylens[:, None])).all()
    assert (2 * ylens <= pred_len(xlens)).all() and (xs[1, int(xlens[1]):] == 0).all()
