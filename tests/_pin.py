"""Shared by the GPU parity tests: a `step_hook` for PPOLearner / ImpalaLearner.update that pins every minibatch step of an
update to the oracle's recorded state (oracle.*Learner.update(record=...))."""
import numpy as np


def pin_hook(record, diag, algo, grad_bar=1e-3, param_bar=2e-6, shard=None, on_last=None):
    """BEFORE minibatch step k the replica is put into the oracle's recorded pre-step state (parameters, optimizer moments,
    count), so every step of the update is a SINGLE-step comparison held to the single-step bars: loss scalars 1e-4 relative,
    gradient `grad_bar` (1e-3), parameters after the optimizer step `param_bar` absolute (2e-6 = 1% of one lr-sized Adam step).
    Chained optimizer steps amplify 1e-6 forward noise through relu / max-pool gate flips (DESIGN.md "chaos caveat"); the
    free-running comparisons this replaces needed 1e-2 .. 0.2 bars that detect nothing.
    `record` is a list (one update) or a callable returning the current update's list; `shard` selects the per-replica
    statistics / gradient of a multi-replica record; `on_last(L)` runs after the last step of the update."""
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
            serr = np.abs(st[:4] - want_s[:4]) / np.maximum(np.abs(want_s[:4]), 1e-6)
            g = L.grads.detach().cpu().numpy().astype(np.float64)
            gerr = float(np.linalg.norm(g - want_g) / np.linalg.norm(want_g))
            diag.append(dict(k=k, shard=shard, stats_relerr=float(serr.max()), grad_relerr=gerr))
            assert serr.max() < 1e-4, (k, st, want_s)          # losses: 1e-4 relative
            assert gerr < grad_bar, (k, gerr)
        else:
            perr = float(np.abs(c.get_params().cpu().numpy() - r["params"]).max())
            diag.append(dict(k=k, shard=shard, params_abs=perr))
            assert perr < param_bar, (k, perr)
            if k == len(rec) - 1:
                # hand the oracle's exact post-update parameters on (parameter publish -> identical actor decisions next rollout)
                c.set_params(r["params"])
                if on_last is not None:
                    on_last(L)
    return hook
