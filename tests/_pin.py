"""Shared by the GPU parity tests: a `step_hook` for PPOLearner / ImpalaLearner.update that pins every minibatch step of an
update to the oracle's recorded state (oracle.*Learner.update(record=...))."""
import numpy as np

from oracle import optim as ooptim


def pin_hook(record, diag, algo, grad_bar=1e-3, param_bar=2e-7, shard=None, nshards=1, shared=None, max_norm=None, on_last=None):
    """BEFORE minibatch step k the replica is put into the oracle's recorded pre-step state (parameters, optimizer moments,
    count), so every step of the update is a SINGLE-step comparison held to the single-step bars:
      * loss scalars: 1e-4 relative to the oracle's (scalars below 1% of the largest one are compared at that floor);
      * gradient: `grad_bar` (1e-3) in relative L2 norm against the oracle's fp32 autograd gradient;
      * optimizer: the parameters after the step against the ORACLE's optimizer (optax chain restated in oracle/optim.py)
        applied to the recorded pre-step state and the replicas' OWN gradients: `param_bar` = 2e-7 absolute.  (Feeding the
        oracle's gradient instead would fold the gradients' fp32 rounding noise, amplified without bound by Adam's
        g / (sqrt(v) + eps) on near-cancelling elements, into a bar on the optimizer arithmetic.)
    Chained optimizer steps amplify 1e-6 forward noise through relu / max-pool gate flips (DESIGN.md "chaos caveat"); the
    free-running comparisons this replaces needed 1e-2 .. 0.2 bars that detect nothing.
    `record` is a list (one update) or a callable returning the current update's list; `shard` / `nshards` select the per-replica
    statistics / gradient of a multi-replica record (`shared`: one dict shared by the replicas' hooks);
    `on_last(L)` runs after the last step of the update."""
    shared = {} if shared is None else shared

    def hook(phase, k, L):
        rec = record() if callable(record) else record
        r = rec[k]
        c = L.ctx
        if phase == "pre":
            c.set_params(r["params_before"])
            if algo == "ppo":
                c.set_opt_state(r["m_before"], r["v_before"], r["count_before"])
            else:
                c.set_opt_state(np.zeros_like(r["nu_before"]), r["nu_before"], r["count_before"])
            L.opt_count = r["count_before"]
        elif phase == "grad":
            want_s = r["stats"] if shard is None else r["shard_stats"][shard]
            want_g = r["raw_grad"] if shard is None else r["shard_grads"][shard]
            st = L.stats[k].detach().cpu().numpy().astype(np.float64)
            # relative to each scalar's own magnitude, floored at 1% of the largest of the four: the policy loss of normalised
            # advantages is a mean of +-O(1) terms that cancels to ~1e-3, where its own magnitude is no scale for a 1e-4 bar
            serr = np.abs(st[:4] - want_s[:4]) / np.maximum(np.abs(want_s[:4]), 1e-2 * np.abs(want_s[:4]).max())
            g32 = L.grads.detach().cpu().numpy()
            shared[(id(rec), k, shard or 0)] = g32
            g = g32.astype(np.float64)
            gerr = float(np.linalg.norm(g - want_g) / np.linalg.norm(want_g))
            gerr64 = None
            if gerr >= grad_bar and "grad64" in r:
                # relu / max-pool gates are discontinuous: the fp32 oracle itself rounds a pre-activation within ~1e-7 of zero
                # differently from box to box (its CPU kernels and thread count differ), which moves the gradient of a
                # 16-sample minibatch by a few 1e-3.  The float64 evaluation of the same step is the tie-breaker.
                g64 = r["grad64"](shard)
                gerr64 = float(np.linalg.norm(g - g64) / np.linalg.norm(g64))
            diag.append(dict(k=k, shard=shard, stats_relerr=float(serr.max()), grad_relerr=gerr, grad_relerr_fp64=gerr64))
            assert serr.max() < 1e-4, (k, st, want_s)          # losses: 1e-4 relative
            assert min(gerr, gerr64 if gerr64 is not None else gerr) < grad_bar, (k, gerr, gerr64)
        else:
            gs = [shared[(id(rec), k, s)] for s in range(nshards)]
            g = gs[0] if nshards == 1 else np.mean(np.stack(gs), axis=0, dtype=np.float32)     # lax.pmean
            if algo == "ppo":
                opt = ooptim.Adam(g.size)
                opt.m, opt.v, opt.count = r["m_before"].copy(), r["v_before"].copy(), r["count_before"]
                mn = 0.5 if max_norm is None else max_norm
            else:
                opt = ooptim.RMSPropPyTorchStyle(g.size)
                opt.nu, opt.count = r["nu_before"].copy(), r["count_before"]
                mn = 40.0 if max_norm is None else max_norm
            want_p = opt.step(r["params_before"], ooptim.clip_by_global_norm(g, mn), r["lr"])
            got_p = c.get_params().cpu().numpy()
            perr = float(np.abs(got_p - want_p).max())
            drift = float(np.abs(got_p - r["params"]).max())    # informational: distance to the oracle's own post-step parameters
            diag.append(dict(k=k, shard=shard, params_abs=perr, params_vs_oracle_grad_abs=drift))
            assert perr < param_bar, (k, perr)
            if k == len(rec) - 1:
                # hand the oracle's exact post-update parameters on (parameter publish -> identical actor decisions next rollout)
                c.set_params(r["params"])
                if on_last is not None:
                    on_last(L)
    return hook
