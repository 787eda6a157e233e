"""Golden-vector case list shared by make_golden.py (which runs every case through
the UNMODIFIED reference) and the parity tests (which replay them against the
oracle port and the CUDA path).  Inputs are seeded recipes, never stored."""

from __future__ import annotations

import numpy as np

from covalent_ssh_plugin_b200 import electrons as E

I64MAX = (1 << 63) - 1
I64MIN = -(1 << 63)


def f32(seed, shape, order="C", dist="normal", view=None):
    return ("f32", seed, tuple(shape) if not isinstance(shape, int) else (shape,), order, dist, view)


# IEEE-754 corner values as bit patterns: signed zeros, the smallest / largest subnormals, the smallest
# normal, values whose 2x+1 or x/2 is a rounding tie, the largest finite, infinities (and, for pure copies,
# NaNs with payloads).  A device that flushes subnormals, fuses 2x+1 into one rounding or canonicalises
# copied NaNs produces different bytes than the reference's numpy on the CPU.
_SPECIAL_BITS = [0x00000000, 0x80000000, 0x00000001, 0x80000001, 0x00000003, 0x007FFFFF, 0x807FFFFF, 0x00800000,
                 0x80800000, 0x00800001, 0x3F800000, 0xBF800000, 0x3F800001, 0x3F7FFFFF, 0x4B000001, 0x4B7FFFFF,
                 0xCB000001, 0x4B800001, 0x33800000, 0x33800001, 0x337FFFFF, 0xB3800000, 0x7F7FFFFF, 0xFF7FFFFF,
                 0x7F000000, 0x7EFFFFFF, 0x7F800000, 0xFF800000, 0x3EAAAAAB, 0x40490FDB, 0x00000002, 0x80000002]
_NAN_BITS = [0x7FC00000, 0xFFC00000, 0x7FC12345, 0x7F800001, 0xFFABCDEF]


def _specials(rng, n, with_nan=False):
    bits = list(_SPECIAL_BITS) + (list(_NAN_BITS) if with_nan else [])
    out = np.empty(n, dtype=np.uint32)
    table = np.array(bits, dtype=np.uint32)
    out[:] = table[np.arange(n) % len(table)]
    # every third element: a random subnormal or a value a few ulps from a power of two
    k = np.arange(n) % 3 == 2
    rnd = rng.integers(0, 1 << 23, size=n, dtype=np.uint32)
    sign = rng.integers(0, 2, size=n, dtype=np.uint32) << np.uint32(31)
    sub = sign | rnd                                                  # exponent 0: subnormals
    near = (np.uint32(0x3F800000) + (rnd & np.uint32(7)) - np.uint32(3)) | sign
    out[k] = np.where(rnd[k] & np.uint32(8), sub[k], near[k])
    return out.view(np.float32)


def build_arg(recipe):
    """Materialise one argument from its recipe."""
    if recipe[0] == "int":
        return recipe[1]
    if recipe[0] == "f32":
        _, seed, shape, order, dist, view = recipe
        rng = np.random.default_rng([0, seed])
        full = tuple(2 * s for s in shape) if view == "stride2" else shape
        n = int(np.prod(full, dtype=np.int64))
        if dist in ("specials", "specials_nan"):
            flat = _specials(rng, n, with_nan=dist == "specials_nan")
        else:
            flat = rng.standard_normal(n, dtype=np.float32) if dist == "normal" else rng.random(n, dtype=np.float32)
        arr = flat.reshape(full, order="C")
        if order == "F":
            arr = np.asfortranarray(arr)
        if view == "stride2":
            arr = arr[tuple(slice(None, None, 2) for _ in shape)]
        return arr
    if recipe[0] == "f32list":
        _, seed, k, shape = recipe
        return [build_arg(f32(seed * 1000 + r, shape, dist="uniform")) for r in range(k)]
    raise ValueError(recipe)


def raise_value_error(tag, n):
    raise ValueError(tag, n)


def plain_python(a, b):          # an UNREGISTERED electron (host semantics only)
    return ", ".join([a, b])


ELECTRONS = {
    "add_int": E.add_int, "mul_int": E.mul_int, "affine_int": E.affine_int, "identity": E.identity,
    "scale_add": E.scale_add, "plus_one": E.plus_one, "vec_add": E.vec_add, "vec_mul": E.vec_mul,
    "axpy_half": E.axpy_half, "vec_sum": E.vec_sum, "stack_sum": E.stack_sum,
    "raise_value_error": raise_value_error, "plain_python": plain_python,
}

_I = lambda v: ("int", v)  # noqa: E731

CASES = [
    # ---- integer electrons (BASELINE config 1/2) -------------------------------
    ("int_add_3_4", "add_int", [_I(3), _I(4)]),
    ("int_add_zero", "add_int", [_I(0), _I(0)]),
    ("int_add_neg", "add_int", [_I(-5), _I(5)]),
    ("int_add_big_ok", "add_int", [_I(1 << 62), _I((1 << 62) - 1)]),
    ("int_add_max_0", "add_int", [_I(I64MAX), _I(0)]),
    ("int_add_overflow_pos", "add_int", [_I(I64MAX), _I(1)]),
    ("int_add_overflow_neg", "add_int", [_I(I64MIN), _I(-1)]),
    ("int_add_beyond_i64", "add_int", [_I(10 ** 30), _I(1)]),
    ("int_mul_7_6", "mul_int", [_I(7), _I(6)]),
    ("int_mul_neg", "mul_int", [_I(-3), _I(1 << 40)]),
    ("int_mul_edge_ok", "mul_int", [_I(1 << 31), _I((1 << 31) - 1)]),
    ("int_mul_overflow", "mul_int", [_I(1 << 32), _I(1 << 32)]),
    ("int_mul_min_neg1", "mul_int", [_I(I64MIN), _I(-1)]),
    ("int_affine_0", "affine_int", [_I(0)]),
    ("int_affine_9999", "affine_int", [_I(9999)]),
    ("int_affine_neg", "affine_int", [_I(-17)]),
    ("int_affine_overflow", "affine_int", [_I(1 << 62)]),
    # ---- fp32 array electrons (BASELINE config 3/4/5) ---------------------------
    ("scale_add_n1", "scale_add", [f32(1, 1)]),
    ("scale_add_n3", "scale_add", [f32(2, 3)]),
    ("scale_add_n4", "scale_add", [f32(3, 4)]),
    ("scale_add_n5", "scale_add", [f32(4, 5)]),
    ("scale_add_n255", "scale_add", [f32(5, 255)]),
    ("scale_add_n1000", "scale_add", [f32(6, 1000)]),
    ("scale_add_64k", "scale_add", [f32(7, 16384)]),
    ("scale_add_1m", "scale_add", [f32(8, 262144)]),
    ("scale_add_empty", "scale_add", [f32(9, 0)]),
    ("scale_add_2d", "scale_add", [f32(10, (128, 96))]),
    ("scale_add_forder", "scale_add", [f32(11, (64, 32), order="F")]),
    ("scale_add_strided", "scale_add", [f32(12, (50, 20), view="stride2")]),
    ("plus_one_64k", "plus_one", [f32(13, 16384)]),
    ("identity_64k", "identity", [f32(14, 16384)]),
    ("identity_3x5", "identity", [f32(15, (3, 5))]),
    ("vec_add_n5", "vec_add", [f32(16, 5), f32(17, 5)]),
    ("vec_add_64k", "vec_add", [f32(18, 16384), f32(19, 16384)]),
    ("vec_mul_64k", "vec_mul", [f32(20, 16384), f32(21, 16384)]),
    ("vec_mul_2d", "vec_mul", [f32(22, (64, 64)), f32(23, (64, 64))]),
    ("axpy_half_64k", "axpy_half", [f32(24, 16384), f32(25, 16384)]),
    ("vec_sum_n1", "vec_sum", [f32(26, 1, dist="uniform")]),
    ("vec_sum_n7", "vec_sum", [f32(27, 7, dist="uniform")]),
    ("vec_sum_n1000", "vec_sum", [f32(28, 1000, dist="uniform")]),
    ("vec_sum_64k", "vec_sum", [f32(29, 16384, dist="uniform")]),
    ("vec_sum_1m", "vec_sum", [f32(30, 262144, dist="uniform")]),
    ("stack_sum_3x5", "stack_sum", [("f32list", 31, 3, (5,))]),
    ("stack_sum_17x1000", "stack_sum", [("f32list", 32, 17, (1000,))]),
    ("stack_sum_64x4096", "stack_sum", [("f32list", 33, 64, (4096,))]),
    # ---- IEEE-754 corner values (subnormals, signed zeros, ties, overflow to inf; NaNs for copies) ----
    ("identity_specials_nan", "identity", [f32(40, 4099, dist="specials_nan")]),
    ("scale_add_specials", "scale_add", [f32(41, 4099, dist="specials")]),
    ("plus_one_specials", "plus_one", [f32(42, 1031, dist="specials")]),
    ("vec_add_specials", "vec_add", [f32(43, 4099, dist="specials"), f32(44, 4099)]),
    ("vec_mul_specials", "vec_mul", [f32(45, 4099, dist="specials"), f32(46, 4099)]),
    ("axpy_half_specials", "axpy_half", [f32(47, 4099, dist="specials"), f32(48, 4099, dist="specials")]),
    # ---- behaviours that never reach the device ---------------------------------
    ("raises_value_error", "raise_value_error", [("lit", "boom"), ("lit", 7)]),
    ("plain_python_join", "plain_python", [("lit", "Hello"), ("lit", "World")]),
]


def build_args(recipes):
    return [r[1] if r[0] == "lit" else build_arg(r) for r in recipes]
