"""The random-number contract of device-generated runs of moves.

The reference draws its random numbers from two sequential Mersenne-Twister streams (Python's ``random`` for the
selector's ``randint``, the translation amplitude and the acceptance test, Core/Collection.py:10-11, :698,
Selectors/RandomSelectors.py:94, Engine.py:3311; ``np.random`` for the direction, Core/Collection.py:686), and the
acceptance number is drawn from the SAME stream as the next step's group index -- only when the move made things worse.
How far the stream has advanced when step j is generated therefore depends on the fate of every earlier step: a run of
moves cannot be generated ahead of its decisions from those streams.

The contract here is COUNTER BASED instead: everything random about step number ``c`` of a run is a pure function of
``(seed, c)``, so the device can generate any step at any time, and a reference Engine equipped with the selector /
generator plug-ins of :mod:`fullrmc_b200.engine_plugins` (its own extension points) draws exactly the same numbers on
the host.  Philox4x32-10 (Salmon et al., SC'11), key = the two halves of the 64-bit seed, counter =
``(c mod 2^32, c >> 32, block, 0)``:

* block 0: word 0 -> group index ``(w0 * numberOfGroups) >> 32``; words 1-3 -> direction, ``v_k = 1 - 2 u_k``;
* block 1: word 0 -> amplitude, word 1 -> the acceptance number ``generate_random_float()`` of this step,
  word 2 -> which generator of a collector (unused by the translation plug-in), word 3 spare;

``u = (w >> 8) * 2^-24`` (a float32 in [0, 1), exact).  The translation vector follows generate_random_vector
(Core/Collection.py:674-701) in float32, one IEEE operation at a time (``translation_vector`` below is the definition;
csrc/rng.cuh repeats it with ``__fmul_rn`` / ``__fadd_rn`` / ``__fdiv_rn`` / ``__fsqrt_rn``), and the moved box
coordinates are ``transform_coordinates(reciprocalBasisVectors, real + vector)``
(Extensions/boundary_conditions_collection.pyx:88-110, Engine.py:3222-3223).
"""
import numpy as np

_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF
F32 = np.float32


def philox4x32(counter, key, rounds=10):
    """counter: 4 uint32, key: 2 uint32 -> 4 uint32 (Random123's philox4x32_R)"""
    c0, c1, c2, c3 = [int(x) & _MASK for x in counter]
    k0, k1 = [int(x) & _MASK for x in key]
    for _ in range(rounds):
        p0 = _M0 * c0
        p1 = _M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _MASK, p1 & _MASK, ((p0 >> 32) ^ c3 ^ k1) & _MASK, p0 & _MASK
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def step_words(seed, counter):
    """the eight 32-bit words of step `counter` (blocks 0 and 1)"""
    key = (seed & _MASK, (seed >> 32) & _MASK)
    lo, hi = counter & _MASK, (counter >> 32) & _MASK
    return philox4x32((lo, hi, 0, 0), key) + philox4x32((lo, hi, 1, 0), key)


def uniform(word):
    """(w >> 8) * 2^-24 as float32"""
    return F32(word >> 8) * F32(2.0 ** -24)


def group_index(word, number_of_groups):
    return (int(word) * int(number_of_groups)) >> 32


def translation_vector(words, min_amp, max_amp):
    """float32 (3,) translation of a step from its words (generate_random_vector, Core/Collection.py:674-701)"""
    min_amp, max_amp = F32(min_amp), F32(max_amp)
    v = [F32(1.0) - F32(2.0) * uniform(words[1 + k]) for k in range(3)]
    n2 = (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]
    if n2 == F32(0.0):
        v, n2 = [F32(1.0), F32(0.0), F32(0.0)], F32(1.0)
    norm = np.sqrt(n2)
    v = [x / norm for x in v]
    amp = uniform(words[4]) * (max_amp - min_amp)
    return np.array([x * amp + x * min_amp for x in v], dtype=F32)


def acceptance_number(words):
    return uniform(words[5])


def transform_coordinates(trans_matrix, coords):
    """float32 restatement of boundary_conditions_collection.transform_coordinates (pyx:88-110), operation by operation"""
    m = np.asarray(trans_matrix, dtype=F32)
    c = np.asarray(coords, dtype=F32)
    out = np.empty_like(c)
    for k in range(3):
        out[:, k] = (c[:, 0] * m[0, k] + c[:, 1] * m[1, k]) + c[:, 2] * m[2, k]
    return out


def generate_step(seed, counter, group_offsets, group_indexes, real_coordinates, reciprocal_basis, min_amp, max_amp):
    """Step `counter` of a run on the current real coordinates: (group, atom indexes, moved real coordinates, moved box
    coordinates, acceptance number).  reciprocal_basis None: non-periodic system, box coordinates are the real ones."""
    w = step_words(seed, counter)
    g = group_index(w[0], len(group_offsets) - 1)
    idx = np.asarray(group_indexes[group_offsets[g]:group_offsets[g + 1]], dtype=np.int32)
    vec = translation_vector(w, min_amp, max_amp)
    moved_real = (np.asarray(real_coordinates, dtype=F32)[idx] + vec).astype(F32)
    moved_box = moved_real if reciprocal_basis is None else transform_coordinates(reciprocal_basis, moved_real)
    return g, idx, moved_real, moved_box, acceptance_number(w)
