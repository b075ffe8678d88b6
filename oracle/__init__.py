"""CPU oracle for the Sebulba hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This package is a CPU restatement (numpy for integer/byte work, PyTorch-CPU fp32/fp64 for the
floating-point network and autograd) of the algorithm behind the two hot paths of
vwxyzjn/cleanba (`cleanba/cleanba_ppo.py`, `cleanba/cleanba_impala.py`).  Every function cites the
reference `file:line` it follows.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  Nothing under
`cleanba_b200/` imports it: the product path is CUDA-only and fails loudly without its extension.

PARITY PINNING STATUS
---------------------
* `oracle.threefry` (JAX 0.4.8 threefry2x32 PRNG: PRNGKey/split/random_bits/uniform/permutation) is
  PINNED: it reproduces the three Random123 threefry2x32-20 known-answer vectors and the public JAX
  documentation constants for `split(PRNGKey(0))` and `uniform(PRNGKey(0))` (tests/test_oracle_prng.py).
* The reference's OWN LINES of the path are PINNED BY EXECUTION: tests/golden/make_reference_exec.py lifts the hot-path
  functions out of cleanba/cleanba_ppo.py and cleanba/cleanba_impala.py with `ast` (get_action_and_value, get_action,
  compute_gae_once / compute_gae, the advantage normalisation block, get_logprob_entropy_value, ppo_loss, impala_loss and its
  two loss wrappers, scale_by_rms_pytorch_style, both linear_schedules, BOTH whole single_device_update functions (on one and on two emulated learner devices), both whole rollout() thread functions, both whole `__main__` blocks, `class Args`
  and the size derivation of `__main__`) and executes those bodies unchanged over PyTorch-CPU stand-ins for the third-party
  names they call.  tests/test_reference_exec.py holds the oracle to the resulting vectors (GAE bit-exact; fp64 loss values
  1e-12 and gradients 1e-9; whole PPO / IMPALA updates: key, optimizer count, scalars 2e-5, parameters) and, under `-m gpu`,
  the CUDA actor step / GAE / PPO update directly.  The stand-ins are restatements (threefry: pinned below; flax modules:
  oracle.network; optax: oracle.optim; rlax: restated a second time in the generator), so the third-party arithmetic itself
  stays in the next category.
* The third-party arithmetic is "PARITY UNPINNED": the reference ships no tests, golden vectors or fixtures
  (SURVEY.md section 4), and its third-party arithmetic (jax 0.4.8, flax 0.6.8, optax 0.1.4, rlax 0.1.5 --
  poetry.lock) is neither vendored under /root/reference nor installable in this image, so the
  reference cannot be run to generate fixtures.  Those parts restate the published semantics of the
  pinned upstream versions and are anchored by closed-form identities and autograd checks
  (tests/test_oracle_*.py).  The golden vectors in tests/golden/ are produced BY THIS ORACLE
  (tests/golden/make_golden.py) and pin it against regressions, not against JAX.
* tests/test_oracle_crosscheck.py compares each restated third-party semantic with an INDEPENDENT implementation or
  derivation (torch.optim.RMSprop / Adam, the IMPALA paper's closed-form V-trace, explicit GAE sums, an explicit conv
  loop, torch.distributions.Categorical, a chi-square test of the Gumbel-max sampler).
* `oracle.carriers` restates the tensor-core operand carriers (bf16 x 3, bf16 x 2, fp16 x 2) bit for bit;
  tests/test_oracle_carriers.py states the accuracy each one guarantees.
"""
