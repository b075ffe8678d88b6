"""Pins oracle.threefry against Random123 KATs and public JAX constants."""
import numpy as np

from oracle import threefry as tf


def test_random123_kats():
    cases = [((0, 0), (0, 0), (0x6B200159, 0x99BA4EFE)),
             ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
             ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]
    for key, ctr, want in cases:
        y0, y1 = tf.threefry2x32_block(key[0], key[1], [ctr[0]], [ctr[1]])
        assert (int(y0[0]), int(y1[0])) == want


def test_jax_doc_constants():
    assert tf.split(tf.PRNGKey(0)).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert np.float32(tf.uniform(tf.PRNGKey(0))) == np.float32(0.41845703)


def test_workload_keys():
    k = tf.split(tf.PRNGKey(1), 4)
    assert k.tolist() == [[869452973, 4133157646], [261504626, 4112007671], [3597360905, 253918841], [98387565, 678776088]]
    nk, sk = tf.split(k[0])
    assert nk.tolist() == [3104884793, 4217725495] and sk.tolist() == [1558003295, 912205106]
    u = tf.uniform(sk, (60, 18)).ravel()[:4]
    np.testing.assert_array_equal(u, np.array([0.7724601, 0.4020629, 0.4750601, 0.61377704], np.float32))


def test_uniform_range_and_odd_sizes():
    u = tf.uniform(tf.PRNGKey(7), (7, 3))  # odd size exercises the pad path
    assert u.dtype == np.float32 and (u >= 0).all() and (u < 1).all()
    assert tf.random_bits(tf.PRNGKey(7), (5,)).shape == (5,)
    assert tf.random_bits(tf.PRNGKey(7), (0,)).shape == (0,)


def test_permutation_properties():
    for n in (1, 2, 512, 2048, 15360):
        assert tf.permutation_rounds(n) == (1 if n <= 1 else 2) or n < 1626
        p = tf.permutation(tf.PRNGKey(3), n)
        assert sorted(p.tolist()) == list(range(n))
    assert tf.permutation_rounds(2048) == 2 and tf.permutation_rounds(15360) == 2 and tf.permutation_rounds(5120) == 2
    # deterministic, key-dependent
    a, b = tf.permutation(tf.PRNGKey(3), 100), tf.permutation(tf.PRNGKey(4), 100)
    assert (a == tf.permutation(tf.PRNGKey(3), 100)).all() and (a != b).any()
