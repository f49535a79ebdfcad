"""Curve and field constants (public parameters of BN254 and BLS12-381).

Oracle / test infrastructure only (see oracle/__init__.py).

The reference obtains these from py_ecc (`curve_order`, `G1`, `field_modulus`,
kzg.py:27-35).  The values below are the published parameters of the two
curves; tests/test_oracle.py checks primality-independent sanity (generator on
curve, r*G1 = O, 2-adicity, primitive roots).
"""

BN254 = {
    "name": "bn254",
    "curve_id": 0,
    # base field modulus p, scalar field modulus (curve order) r
    "p": 21888242871839275222246405745257275088696311157297823662689037894645226208583,
    "r": 21888242871839275222246405745257275088548364400416034343698204186575808495617,
    "b": 3,
    "G1": (1, 2),
    "fr_generator": 5,     # least primitive root mod r (hard part 7)
    "fr_two_adicity": 28,
    "fp_limbs32": 8,
    "fr_limbs32": 8,
}

BLS12_381 = {
    "name": "bls12_381",
    "curve_id": 1,
    "p": 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB,
    "r": 52435875175126190479447740508185965837690552500527637822603658699938581184513,
    "b": 4,
    "G1": (
        0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
        0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
    ),
    "fr_generator": 7,
    "fr_two_adicity": 32,
    "fp_limbs32": 12,
    "fr_limbs32": 8,
}

CURVES = {"bn254": BN254, "bls12_381": BLS12_381}


def curve(name):
    """Same error convention as reference kzg.py:36-37."""
    try:
        return CURVES[name]
    except KeyError:
        raise ValueError(f"Unsupported curve type: {name}")


def root_of_unity(cv, n):
    """w = g^((r-1)/n): what Sage's Fq(1).nth_root(n) yields for the callers
    (plonk/encoder.py:49, marlin/encoder.py:48-49; SURVEY.md section 7 hard part 7)."""
    r = cv["r"]
    assert n & (n - 1) == 0 and n >= 1
    assert (r - 1) % n == 0, "n exceeds the 2-adicity of r-1"
    return pow(cv["fr_generator"], (r - 1) // n, r)
