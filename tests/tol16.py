"""Tolerance policy of the 16-bit tensor-core path (tests only).

BASELINE.json north_star: max-abs mel error <= 2e-2 for the 16-bit path.  The product's 16-bit path - the one bench.py
measures and DESIGN.md recommends - uses FP16 MMA operands (same tcgen05 rate as bf16, 11 mantissa bits): for it the bound is
asserted ABSOLUTE, on every config, with no scale-dependent escape.

``b2s_precision='bf16'`` stays available as an operand mode (wider exponent range for checkpoints whose activations leave the
fp16 range).  With random-init from-noise sampling |mel| reaches ~300 and one bf16 ulp of the layer input alone is 0.5, so its
error is REPORTED (gpurun_out/parity_report.jsonl) next to the fp16 figure; the only assertion on it is a regression guard
relative to |mel|max - it is NOT a claim that bf16 meets the north-star bound."""

TC_TOL = 2e-2             # the stated bound, absolute, asserted for fp16
BF16_GUARD_ABS, BF16_GUARD_REL = 5e-2, 4e-4     # regression guard for the bf16 operand mode: error <= max(5e-2, 4e-4 * |mel|max)


def check16(precision, err, scale, what=''):
    if precision == 'fp16':
        assert err <= TC_TOL, (what, precision, err, scale)
    elif precision == 'bf16':
        assert err <= max(BF16_GUARD_ABS, BF16_GUARD_REL * scale), (what, precision, err, scale, 'bf16 regression guard')
    else:
        raise AssertionError(f'unknown 16-bit precision {precision!r}')
