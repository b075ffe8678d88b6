"""Golden vectors made by EXECUTING THE REFERENCE'S OWN SOURCE LINES for the hot path.

The reference (vwxyzjn/cleanba) is JAX code and jax / flax / optax / rlax / envpool / tyro are not installable in this image,
so `cleanba/cleanba_ppo.py` cannot be imported.  What CAN be done is narrower and still worth having: this script parses the
reference files (read-only, /root/reference), lifts the hot-path functions out of them by name with `ast` -- the function
bodies are the reference's, untouched except that decorators and type annotations are dropped -- and executes them in a
namespace where the THIRD-PARTY names they call (`jnp.*`, `jax.lax.scan`, `jax.nn.log_softmax`, `jax.random.*`, `jax.vmap`,
`jax.value_and_grad`, `rlax.*`, flax modules, the optax chain) are bound to small PyTorch-CPU stand-ins.  The outputs are
committed as tests/golden/reference_exec.npz; tests/test_reference_exec.py compares the oracle (and, under `-m gpu`, the CUDA
path) with them.

What this pins: every line the REFERENCE ITSELF wrote for the path -- GAE recurrence and operand order, bootstrap concat,
advantage normalisation axes, log-prob / entropy formula, PPO clipping and loss assembly, the epoch shuffle (key split, permutation
of the flattened batch, minibatch reshape), the scan order of minibatches and epochs, the averaging of the reported scalars, the same update on two
emulated learner devices (pmean'ed gradients, local advantage normalisation, shared shuffle key), the whole rollout() thread loop, the
IMPALA slicing ([:-1] / [1:]), discount / mask construction, the T-scaled sums of the rlax losses, the contiguous column split of
IMPALA minibatches, the RMSProp-pytorch-style update rule, both learning-rate schedules, the Gumbel-max sampling line, the Args
defaults and the size derivation of `__main__`.
What it does NOT pin: the third-party arithmetic behind those names.  The stand-ins restate it (threefry / permutation:
oracle.threefry, itself pinned by known-answer vectors; flax modules: oracle.network; optax clip / adam / MultiSteps(k=1):
oracle.optim; rlax 0.1.5 vtrace / policy_gradient_loss / entropy_loss / importance ratios: restated HERE from the published
source, independently of oracle/impala.py), and XLA's kernel numerics are out of reach.

Run (needs /root/reference; the tests only need the committed .npz):   python tests/golden/make_reference_exec.py
"""
import ast
import collections
import dataclasses
import json
import os
import sys
import types
from functools import partial
from typing import List, NamedTuple, Optional

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import network as net, optim, threefry as tf  # noqa: E402  (third-party stand-ins only)

REF = os.environ.get("CLEANBA_REFERENCE", "/root/reference")
NS = types.SimpleNamespace


# ------------------------------------------------------------------------------------------------ lifting reference code
class _Strip(ast.NodeTransformer):
    """Drops decorators (@jax.jit) and annotations (flax / jax types that do not exist here); bodies are untouched."""

    def visit_FunctionDef(self, node):
        self.generic_visit(node)
        node.decorator_list = []
        node.returns = None
        for a in node.args.posonlyargs + node.args.args + node.args.kwonlyargs:
            a.annotation = None
        return node


def lift(tree, name, kind=(ast.FunctionDef, ast.ClassDef)):
    """Source of the first def / class called `name` anywhere in the module (nested defs included)."""
    for node in ast.walk(tree):
        if isinstance(node, kind) and node.name == name:
            return ast.unparse(ast.fix_missing_locations(_Strip().visit(node)))
    raise KeyError(name)


def lift_stmt(tree, pred):
    for node in ast.walk(tree):
        if isinstance(node, ast.stmt) and pred(node):
            return ast.unparse(node)
    raise KeyError("statement not found")


def run(src, ns):
    exec(compile(src, "<reference>", "exec"), ns)


# ------------------------------------------------------------------------------------------------ jax stand-ins (torch CPU)
class JT(torch.Tensor):
    """Tensor with the two jnp behaviours torch lacks: population std (ddof = 0) and arithmetic on bool arrays."""
    _ARITH = {"sub", "rsub", "__sub__", "__rsub__", "add", "__add__", "__radd__", "mul", "__mul__", "__rmul__"}

    @classmethod
    def __torch_function__(cls, func, types_, args=(), kwargs=None):
        kwargs = dict(kwargs or {})
        name = getattr(func, "__name__", "")
        if name == "std" and "correction" not in kwargs and "unbiased" not in kwargs:
            kwargs["correction"] = 0
        if name in cls._ARITH:
            dt = torch.get_default_dtype()
            args = tuple(a.to(dt) if isinstance(a, torch.Tensor) and a.dtype == torch.bool else a for a in args)
        return super().__torch_function__(func, types_, args, kwargs)

    # jax arrays are immutable: `total_loss += x` (cleanba_impala.py:594-595) rebinds, it must not write through an alias
    def __iadd__(self, o):
        return self + o

    def __isub__(self, o):
        return self - o

    def __imul__(self, o):
        return self * o

    def __itruediv__(self, o):
        return self / o


def frames(seed, shape):
    """Synthetic uint8 frames from a seed (the fixture stores the seed, not megabytes of incompressible bytes)."""
    return np.random.default_rng(int(seed)).integers(0, 256, shape, dtype=np.uint8)


def J(x, dtype=None):
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.as_tensor(np.asarray(x))
    if dtype is not None:
        t = t.to(dtype)
    elif t.dtype in (torch.float32, torch.float64):
        t = t.to(torch.get_default_dtype())
    return t.as_subclass(JT)


def _is_leaf(x):
    return isinstance(x, (torch.Tensor, np.ndarray)) or np.isscalar(x) or x is None


def tree_map(f, tree, *rest):
    if isinstance(tree, tuple) and hasattr(tree, "_fields"):
        return type(tree)(*[tree_map(f, x, *[r[i] for r in rest]) for i, x in enumerate(tree)])
    if isinstance(tree, (list, tuple)):
        return type(tree)(tree_map(f, x, *[r[i] for r in rest]) for i, x in enumerate(tree))
    if isinstance(tree, dict):
        return {k: tree_map(f, v, *[r[k] for r in rest]) for k, v in tree.items()}
    return f(tree, *rest)


def _stack_tree(items, axis=0):
    first = items[0]
    if isinstance(first, tuple) and hasattr(first, "_fields"):
        return type(first)(*[_stack_tree([it[i] for it in items], axis) for i in range(len(first))])
    if isinstance(first, (list, tuple)):
        return type(first)(_stack_tree([it[i] for it in items], axis) for i in range(len(first)))
    return torch.stack([torch.as_tensor(it) for it in items], dim=axis).as_subclass(JT)


def scan(f, init, xs, length=None, reverse=False):
    """jax.lax.scan: carry threaded through f over the leading axis of every leaf of xs; ys stacked."""
    leaves = []
    tree_map(lambda x: leaves.append(x), xs)
    n = length if length is not None else leaves[0].shape[0]
    order = range(n - 1, -1, -1) if reverse else range(n)
    carry, ys = init, [None] * n
    for i in order:
        carry, ys[i] = f(carry, tree_map(lambda x: x[i], xs))
    return carry, (_stack_tree(ys) if n and ys[0] is not None and ys[0] != () else ys[0] if n else None)


def vmap(f, in_axes=0, out_axes=0):
    def g(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        if tuple(axes) == (None, 0) and out_axes == 0 and args[1].ndim >= 3:
            # vmap of a batched function over a leading axis IS batching: one call on the flattened [T*B, ...] frames (as XLA
            # executes it, and with the same CPU conv kernels as a direct batched call -- a per-row loop rounds differently)
            x = args[1]
            out = f(args[0], x.reshape((-1,) + tuple(x.shape[2:])))
            return tree_map(lambda t: t.reshape(tuple(x.shape[:2]) + tuple(t.shape[1:])), out)
        n = next(a.shape[ax] for a, ax in zip(args, axes) if ax is not None)
        outs = [f(*[a if ax is None else a.select(ax, i) for a, ax in zip(args, axes)]) for i in range(n)]
        return _stack_tree(outs, out_axes)
    return g


def _split(x, n, axis=0):
    return [t.as_subclass(JT) for t in torch.tensor_split(x, n, dim=axis)]


def _jnp_array(x, dtype=None):
    if isinstance(x, (list, tuple)) and len(x) and isinstance(x[0], torch.Tensor):
        return torch.stack(list(x)).as_subclass(JT)
    return J(x, dtype)


jnp = NS(
    array=_jnp_array, asarray=_jnp_array, zeros_like=lambda x: torch.zeros_like(x).as_subclass(JT),
    concatenate=lambda xs, axis=0: torch.cat([x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x)) for x in xs], dim=axis).as_subclass(JT),
    hstack=lambda xs: torch.hstack(list(xs)),
    stack=lambda xs, axis=0: torch.stack([torch.as_tensor(np.asarray(x)) if not isinstance(x, torch.Tensor) else x for x in xs], dim=axis).as_subclass(JT),
    split=_split, zeros=lambda n, dtype=None: torch.zeros(n, dtype=dtype).as_subclass(JT),
    exp=torch.exp, log=torch.log, clip=lambda x, a, b: torch.clamp(x, a, b), maximum=torch.maximum, minimum=torch.minimum,
    arange=lambda n: torch.arange(int(n)), finfo=torch.finfo, sum=torch.sum, square=torch.square,
    reshape=lambda x, shape: x.reshape(tuple(shape)), argmax=lambda x, axis=None: torch.argmax(x, dim=axis),
    full_like=lambda x, v: torch.full_like(x, v).as_subclass(JT), bool_=torch.bool, ndarray=torch.Tensor,
)


class _Key(np.ndarray):
    """PRNG keys stay numpy uint32[2]; `key, subkey = jax.random.split(key)` unpacks rows."""


def _permutation(key, x):
    if isinstance(x, torch.Tensor) and x.ndim >= 1:
        return x[torch.as_tensor(tf.permutation(np.asarray(key), x.shape[0]).astype(np.int64))]
    return J(tf.permutation(np.asarray(key), int(x)))


STOP_GRADIENT = lambda x: x.detach()  # noqa: E731


def value_and_grad(f, has_aux=False):
    """jax.value_and_grad w.r.t. the first argument (the AgentParams pytree = ONE flat vector here)."""
    def g(params, *a):
        flat = params.flat.detach().clone().requires_grad_(True)
        out = f(Params(flat), *a)
        loss, aux = out if has_aux else (out, None)
        (grad,) = torch.autograd.grad(loss, flat)
        aux = tree_map(lambda t: t.detach(), aux)
        return ((loss.detach(), aux), grad.detach()) if has_aux else (loss.detach(), grad.detach())
    return g


jax = NS(
    jit=lambda f: f, numpy=jnp, tree_map=tree_map, tree_util=NS(tree_map=tree_map), vmap=vmap, value_and_grad=value_and_grad,
    nn=NS(log_softmax=lambda x: torch.log_softmax(x, -1), softmax=lambda x: torch.softmax(x, -1)),
    scipy=NS(special=NS(logsumexp=lambda x, axis=-1, keepdims=False: torch.logsumexp(x, dim=axis, keepdim=keepdims))),
    lax=NS(scan=scan, stop_gradient=lambda x: STOP_GRADIENT(x), pmean=lambda x, axis_name=None: x, sqrt=torch.sqrt),
    random=NS(split=lambda key, num=2: tf.split(np.asarray(key), num), permutation=_permutation, PRNGKey=tf.PRNGKey,
              uniform=lambda key, shape=(): J(tf.uniform(np.asarray(key), tuple(shape)))),
)


# flax modules (cleanba_ppo.py:140-203): restated by oracle.network; here only the .apply call shape of the reference
class _Tree(dict):
    """Leaves by flat flax path; `p["params"]["Dense_0"]["kernel"].block_until_ready()` (cleanba_ppo.py:299-301) is a no-op."""

    def __missing__(self, key):
        return self

    def block_until_ready(self):
        return self


class Params:
    """AgentParams(network_params, actor_params, critic_params) over one flat vector."""

    def __init__(self, flat):
        self.flat = flat
        self._p = None

    def _tree(self):
        if self._p is None:
            self._p = _Tree(net.unflatten(self.flat))
        return self._p

    network_params = actor_params = critic_params = property(_tree)


class Network:
    def __init__(self, channels=None, hiddens=None):
        pass

    def apply(self, p, obs):
        return net.trunk_forward(p, torch.as_tensor(obs)).as_subclass(JT)


class Actor:
    def __init__(self, action_dim=None):
        pass

    def apply(self, p, hidden):
        return net.heads(p, hidden)[0]


class Critic:
    def apply(self, p, hidden):
        return net.heads(p, hidden)[1][:, None]          # flax Dense(1): [n, 1]


class TrainState:
    """flax TrainState + the optax chain of the script, restated by oracle.optim:
    MultiSteps(k=1) o clip_by_global_norm o inject_hyperparams(adam | rmsprop_pytorch_style)(learning_rate=schedule)."""

    def __init__(self, flat, opt, max_grad_norm, schedule, k=1, mini=0, acc=None):
        self.flat, self.opt, self.max_grad_norm, self.schedule = flat, opt, max_grad_norm, schedule
        self.k, self.mini, self.acc = k, mini, acc

    @property
    def params(self):
        return Params(self.flat)

    def apply_gradients(self, grads):
        if self.k > 1:
            # optax.MultiSteps(every_k_schedule=k), optax 0.1.4: acc += (g - acc) / (mini_step + 1) from zeros; the inner chain
            # (and its schedule count) advances on the k-th mini-step only, the other mini-steps emit zero updates
            g32 = grads.numpy().astype(np.float32)
            acc = g32.copy() if self.mini == 0 else (self.acc + (g32 - self.acc) / np.float32(self.mini + 1)).astype(np.float32)
            if self.mini < self.k - 1:
                return TrainState(self.flat, self.opt, self.max_grad_norm, self.schedule, self.k, self.mini + 1, acc)
            grads = torch.from_numpy(acc)
        g = optim.clip_by_global_norm(grads.numpy().astype(np.float32), self.max_grad_norm)
        lr = np.float32(self.schedule(self.opt.count))          # the REFERENCE's linear_schedule at the pre-increment count
        new = self.opt.step(self.flat.detach().numpy().astype(np.float32), g, lr)
        return TrainState(J(new), self.opt, self.max_grad_norm, self.schedule, self.k)


# rlax 0.1.5 (poetry.lock), restated from its published source -- independently of oracle/impala.py
def _log_prob(logits, a):
    return torch.log_softmax(logits, -1).gather(-1, a.long()[..., None]).squeeze(-1)


def _rlax_ratios(pi_logits_t, mu_logits_t, a_t):
    return torch.exp(_log_prob(pi_logits_t, a_t) - _log_prob(mu_logits_t, a_t))


def _rlax_vtrace(v_tm1, v_t, r_t, discount_t, rho_tm1, lambda_, clip_rho_threshold, stop_target_gradients):
    c_tm1 = torch.clamp(rho_tm1, max=1.0) * lambda_
    clipped_rhos_tm1 = torch.clamp(rho_tm1, max=clip_rho_threshold)
    td_errors = clipped_rhos_tm1 * (r_t + discount_t * v_t - v_tm1)
    err, errors = 0.0, []
    for i in reversed(range(v_t.shape[0])):
        err = td_errors[i] + discount_t[i] * c_tm1[i] * err
        errors.insert(0, err)
    targets_tm1 = torch.stack(errors) + v_tm1
    if stop_target_gradients:
        targets_tm1 = STOP_GRADIENT(targets_tm1)
    return targets_tm1 - v_tm1


VTraceOutput = collections.namedtuple("vtrace_output", ["errors", "pg_advantage", "q_estimate"])


def _rlax_vtrace_td_error_and_advantage(v_tm1, v_t, r_t, discount_t, rho_tm1, lambda_=1.0, clip_rho_threshold=1.0,
                                        clip_pg_rho_threshold=1.0, stop_target_gradients=True):
    lambda_ = torch.ones_like(discount_t) * lambda_
    errors = _rlax_vtrace(v_tm1, v_t, r_t, discount_t, rho_tm1, lambda_, clip_rho_threshold, stop_target_gradients)
    targets_tm1 = errors + v_tm1
    q_bootstrap = torch.cat([lambda_[:-1] * targets_tm1[1:] + (1 - lambda_[:-1]) * v_tm1[1:], v_t[-1:]], 0)
    q_estimate = r_t + discount_t * q_bootstrap
    rho_clipped = torch.clamp(rho_tm1, max=clip_pg_rho_threshold)
    return VTraceOutput(errors=errors, pg_advantage=rho_clipped * (q_estimate - v_tm1), q_estimate=q_estimate)


def _rlax_policy_gradient_loss(logits_t, a_t, adv_t, w_t, use_stop_gradient=True):
    adv_t = STOP_GRADIENT(adv_t) if use_stop_gradient else adv_t
    return torch.mean(-_log_prob(logits_t, a_t) * adv_t * w_t)


def _rlax_entropy_loss(logits_t, w_t):
    lp = torch.log_softmax(logits_t, -1)
    p = torch.softmax(logits_t, -1)
    entropy = -torch.where(p == 0, torch.zeros_like(p), p * lp).sum(-1)       # distrax.Softmax(...).entropy()
    return -torch.mean(entropy * w_t)


rlax = NS(categorical_importance_sampling_ratios=_rlax_ratios, vtrace_td_error_and_advantage=_rlax_vtrace_td_error_and_advantage,
          policy_gradient_loss=_rlax_policy_gradient_loss, entropy_loss=_rlax_entropy_loss)

# optax names used by scale_by_rms_pytorch_style (cleanba_impala.py:152-170)
ScaleByRmsState = collections.namedtuple("ScaleByRmsState", ["nu"])
base = NS(GradientTransformation=collections.namedtuple("GradientTransformation", ["init", "update"]))


def update_moment_per_elem_norm(updates, moments, decay, order):
    """optax 0.1.4: (1 - decay) * |g|^order + decay * t, leaf by leaf."""
    assert order == 2
    return tree_map(lambda g, t: (1 - decay) * (g * g) + decay * t, updates, moments)


# ------------------------------------------------------------------------------------------------ the runs
def base_ns(**extra):
    ns = dict(jax=jax, jnp=jnp, np=np, rlax=rlax, partial=partial, List=List, Optional=Optional, NamedTuple=NamedTuple,
              Network=Network, Actor=Actor, Critic=Critic, network=Network(), critic=Critic(), actor=Actor(),
              envs=NS(single_action_space=NS(n=18)), base=base, ScaleByRmsState=ScaleByRmsState,
              update_moment_per_elem_norm=update_moment_per_elem_norm)
    ns.update(extra)
    return ns


def ref_args(tree, **over):
    """The reference's own `class Args` (dataclass defaults) + the size derivation statements of `__main__`."""
    ns = dict(dataclass=dataclasses.dataclass, field=dataclasses.field, List=List, Optional=Optional, os=NS(path=NS(basename=lambda p: "cleanba"), environ={}),
              __file__="cleanba.py", __name__="ref")
    run(lift(tree, "Args"), ns)
    a = ns["Args"]()
    for k, v in over.items():
        setattr(a, k, v)
    return a


def derive_sizes(tree, a, world_size=1):
    """Executes the Assign / Assert statements of the reference's `if __name__ == "__main__":` block that derive the sizes."""
    main = next(n for n in tree.body if isinstance(n, ast.If) and "__main__" in ast.unparse(n.test))
    ns = dict(args=a, jax=NS(process_count=lambda: world_size, process_index=lambda: 0), int=int, len=len)
    wanted = ("local_batch_size", "local_minibatch_size", "world_size", "local_rank", "num_envs", "batch_size", "minibatch_size", "num_updates")
    for st in main.body:
        if isinstance(st, ast.Assert):
            run(ast.unparse(st), ns)
        elif isinstance(st, ast.Assign) and isinstance(st.targets[0], ast.Attribute) and st.targets[0].attr in wanted:
            run(ast.unparse(st), ns)
        if isinstance(st, ast.Assign) and isinstance(st.targets[0], ast.Attribute) and st.targets[0].attr == "num_updates":
            break
    return {k: int(getattr(a, k)) for k in wanted}


def args_defaults(a):
    return {f.name: getattr(a, f.name) for f in dataclasses.fields(a)}


def t2n(x):
    return x.detach().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def build():
    torch.set_num_threads(1)
    torch.set_default_dtype(torch.float32)
    ppo = ast.parse(open(os.path.join(REF, "cleanba", "cleanba_ppo.py")).read())
    imp = ast.parse(open(os.path.join(REF, "cleanba", "cleanba_impala.py")).read())
    rng = np.random.default_rng(77)
    out = {}

    # ---- Args defaults and size derivation (cleanba_ppo.py:34-118, 409-440; cleanba_impala.py likewise), BASELINE configs
    meta = {"ppo_args": args_defaults(ref_args(ppo)), "impala_args": args_defaults(ref_args(imp)), "sizes": []}
    for script, tree, over in (("ppo", ppo, dict(local_num_envs=60)), ("impala", imp, dict(local_num_envs=60)),
                               ("ppo", ppo, dict(local_num_envs=60, actor_device_ids=[0], learner_device_ids=[1, 2, 3])),
                               ("ppo", ppo, dict(local_num_envs=60, actor_device_ids=[0], learner_device_ids=[1, 2, 3], world_size=2)),
                               ("ppo", ppo, dict()), ("impala", imp, dict())):
        ws = over.pop("world_size", 1)
        a = ref_args(tree, **over)
        meta["sizes"].append(dict(script=script, overrides=over, world_size=ws, derived=derive_sizes(tree, a, ws)))

    # ---- learning-rate schedules (cleanba_ppo.py:475-479, cleanba_impala.py:515-519)
    for name, tree in (("ppo", ppo), ("impala", imp)):
        a = ref_args(tree, local_num_envs=60)
        derive_sizes(tree, a)
        ns = dict(args=a)
        run(lift(tree, "linear_schedule"), ns)
        counts = np.array([0, 1, 3, 4, 15, 16, 17, 1000, 52079, a.num_updates * 4 - 1], np.int64)      # inside the annealing range of both scripts
        out[f"{name}_sched_counts"] = counts
        out[f"{name}_sched_lr"] = np.array([ns["linear_schedule"](int(c)) for c in counts], np.float64)
        out[f"{name}_sched_num_updates"] = np.int64(a.num_updates)

    # ---- actor: get_action_and_value (cleanba_ppo.py:245-261), get_action (cleanba_impala.py:287-301)
    flat = net.init_params(3)
    obs = frames(500, (6, 4, 84, 84))
    key = tf.split(tf.PRNGKey(5), 4)[1]
    a = ref_args(ppo)
    ns = base_ns(args=a)
    run(lift(ppo, "get_action_and_value"), ns)
    with torch.no_grad():
        o, action, logprob, value, key2 = ns["get_action_and_value"](Params(J(flat)), obs, key)
    out.update(act_params_seed=np.int64(3), act_obs_seed=np.int64(500), act_key=key, act_action=t2n(action).astype(np.int32), act_logprob=t2n(logprob),
               act_value=t2n(value), act_key_after=np.asarray(key2))
    ns = base_ns(args=ref_args(imp))
    run(lift(imp, "get_action"), ns)
    with torch.no_grad():
        o, action_i, logits_i, key3 = ns["get_action"](Params(J(flat)), obs, key)
    out.update(act_impala_action=t2n(action_i).astype(np.int32), act_impala_logits=t2n(logits_i), act_impala_key_after=np.asarray(key3))

    # ---- compute_gae_once / compute_gae (cleanba_ppo.py:532-560) + advantage normalisation (592-595), fp32
    T, B = 16, 12
    Transition = collections.namedtuple("Transition", ["obs", "dones", "actions", "logprobs", "values", "env_ids", "rewards", "truncations", "terminations", "firststeps"])
    rewards = rng.choice([-1.0, 0.0, 1.0, 0.5], size=(T, B)).astype(np.float32)
    values = (rng.standard_normal((T, B)) * 0.7).astype(np.float32)
    dones = rng.random((T, B)) < 0.15
    next_value = (rng.standard_normal(B) * 0.7).astype(np.float32)
    next_done = rng.random(B) < 0.25

    class FixedCritic:           # compute_gae evaluates the bootstrap value through the network; here a given vector
        def apply(self, p, hidden):
            return J(next_value)[:, None]
    a = ref_args(ppo)
    ns = base_ns(args=a, critic=FixedCritic(), network=NS(apply=lambda p, o: o))
    run(lift(ppo, "compute_gae_once"), ns)
    run(lift_stmt(ppo, lambda n: isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") == "compute_gae_once"), ns)
    run(lift(ppo, "compute_gae"), ns)
    storage = Transition(obs=None, dones=J(dones), actions=None, logprobs=None, values=J(values), env_ids=None, rewards=J(rewards),
                         truncations=None, terminations=None, firststeps=None)
    adv, ret = ns["compute_gae"](NS(params=NS(critic_params=None, network_params=None)), J(np.zeros(B, np.float32)), J(next_done), storage)
    norm_src = lift_stmt(ppo, lambda n: isinstance(n, ast.If) and ast.unparse(n.test) == "args.norm_adv")
    ns2 = dict(args=NS(norm_adv=True, num_minibatches=4), advantages=adv)
    run(norm_src, ns2)
    out.update(gae_rewards=rewards, gae_values=values, gae_dones=dones, gae_next_value=next_value, gae_next_done=next_done,
               gae_adv=t2n(adv), gae_ret=t2n(ret), gae_norm4=t2n(ns2["advantages"]))

    # ---- get_logprob_entropy_value + ppo_loss (cleanba_ppo.py:516-530, 562-577) given the head outputs; value + gradient (fp64)
    torch.set_default_dtype(torch.float64)
    n = 24
    logits0 = rng.standard_normal((n, 18)) * 1.5
    logits0[0, 3] = -800.0                                   # a vanishing probability (the clip(min=finfo.min) line)
    value0 = rng.standard_normal(n)
    acts = rng.integers(0, 18, n).astype(np.int64)
    blp = np.log(rng.dirichlet(np.ones(18), n))[np.arange(n), acts] * 0.9
    advs = rng.standard_normal(n)
    tgt = rng.standard_normal(n)

    class HeadsGiven:                                        # the network's outputs as differentiable inputs
        def __init__(self, t):
            self.t = t

        def __call__(self, *a, **k):
            return self

        def apply(self, p, x):
            return self.t
    lg = J(logits0).requires_grad_(True)
    vl = J(value0).requires_grad_(True)
    ns = base_ns(args=ref_args(ppo), Network=HeadsGiven(None), Actor=HeadsGiven(lg), Critic=HeadsGiven(vl[:, None]))
    run(lift(ppo, "get_logprob_entropy_value"), ns)
    run(lift(ppo, "ppo_loss"), ns)
    lp, ent, val = ns["get_logprob_entropy_value"](NS(network_params=None, actor_params=None, critic_params=None), None, torch.as_tensor(acts))
    loss, (pg, vloss, eloss, kl) = ns["ppo_loss"](NS(network_params=None, actor_params=None, critic_params=None), None, torch.as_tensor(acts),
                                                  J(blp), None, J(advs), J(tgt))
    dlg, dvl = torch.autograd.grad(loss, (lg, vl))
    out.update(ppo_logits=logits0, ppo_value=value0, ppo_actions=acts, ppo_behavior_logprobs=blp, ppo_advantages=advs, ppo_targets=tgt,
               ppo_newlogprob=t2n(lp), ppo_entropy=t2n(ent), ppo_newvalue=t2n(val),
               ppo_scalars=np.array([loss.item(), pg.item(), vloss.item(), eloss.item(), kl.item()]), ppo_dlogits=t2n(dlg), ppo_dvalue=t2n(dvl))

    # ---- impala_loss (cleanba_impala.py:557-597) given the head outputs; value + gradient (fp64)
    T1, Bc = 9, 5
    pl0 = rng.standard_normal((T1, Bc, 18)) * 1.2
    nv0 = rng.standard_normal((T1, Bc))
    bl0 = pl0 + rng.standard_normal((T1, Bc, 18)) * 0.4
    ai = rng.integers(0, 18, (T1, Bc)).astype(np.int64)
    ri = rng.choice([-1.0, 0.0, 1.0], size=(T1, Bc))
    di = rng.random((T1, Bc)) < 0.2
    fi = rng.random((T1, Bc)) < 0.15
    plg = J(pl0).requires_grad_(True)
    nvg = J(nv0).requires_grad_(True)
    ns = base_ns(args=ref_args(imp))
    for fn in ("policy_gradient_loss", "entropy_loss_fn", "impala_loss"):
        run(lift(imp, fn), ns)
    ns["get_logits_and_value"] = None
    # the reference evaluates the network as jax.vmap(get_logits_and_value, (None, 0))(params, x): [T+1, B] head outputs
    ns["jax"] = NS(**{**jax.__dict__, "vmap": lambda f, in_axes=0, out_axes=0: (lambda params, x: (plg, nvg)) if f is None else vmap(f, in_axes, out_axes)})
    total, (pgl, bll, enl) = ns["impala_loss"](None, None, torch.as_tensor(ai), J(bl0), J(ri), J(di), J(fi))
    dpl, dnv = torch.autograd.grad(total, (plg, nvg))
    out.update(imp_policy_logits=pl0, imp_values=nv0, imp_behaviour_logits=bl0, imp_actions=ai, imp_rewards=ri, imp_dones=di, imp_firststeps=fi,
               imp_scalars=np.array([total.item(), pgl.item(), bll.item(), enl.item()]), imp_dlogits=t2n(dpl), imp_dvalue=t2n(dnv))

    # ---- scale_by_rms_pytorch_style (cleanba_impala.py:152-170): three updates of the reference's own transform
    ns = base_ns()
    run(lift(imp, "scale_by_rms_pytorch_style"), ns)
    tx = ns["scale_by_rms_pytorch_style"](decay=0.99, eps=0.01)
    g_seq = rng.standard_normal((3, 40)) * np.array([1.0, 0.1, 10.0])[:, None]
    state = tx.init(J(np.zeros(40)))
    ups = []
    for g in g_seq:
        u, state = tx.update(J(g), state)
        ups.append(t2n(u))
    out.update(rms_grads=g_seq, rms_updates=np.stack(ups), rms_nu=t2n(state.nu))
    torch.set_default_dtype(torch.float32)

    # ---- the whole single_device_update of both scripts (cleanba_ppo.py:579-654, cleanba_impala.py:599-639), fp32, tiny shapes
    out.update(run_ppo_update(ppo, rng))
    out.update(run_impala_update(imp, rng))
    out.update(run_rollout(ppo, "ppo", rng))
    out.update(run_rollout(imp, "impala", rng))
    out.update(run_benchmark_tool())
    out.update(run_main(ppo, "ppo"))
    out.update(run_main(imp, "impala"))
    # the PPO script with --concurrency (rollout u+1 beside update u, the `update != 2` rule of cleanba_ppo.py:287-304), three updates
    out.update(run_main(ppo, "ppo", tag="ppoconc", extra=dict(concurrency=True, total_timesteps=3 * 4 * 3 * 2)))
    # two learner devices: flax replicate, device_put_sharded of the env-axis halves, pmap over both, unreplicate for the actors
    out.update(run_main(ppo, "ppo", tag="ppol2", extra=dict(learner_device_ids=[0, 1])))
    out.update(run_main(imp, "impala", tag="impalal2", extra=dict(learner_device_ids=[0, 1])))
    out["meta_json"] = np.array(json.dumps(meta))
    return out


def digest(prefix, before, after):
    """The updated parameter vector as a strided sample + the norm of the whole step (4.4 MB would not be a small fixture)."""
    d = after.astype(np.float64) - before.astype(np.float64)
    return {f"{prefix}_params_after_every211": after[::211].copy(), f"{prefix}_step_l2": np.float64(np.sqrt((d * d).sum())),
            f"{prefix}_step_sum": np.float64(d.sum())}


class PMean:
    """jax.lax.pmean over L emulated devices: every device runs the reference's update function in its own thread and the
    threads meet here (same call sequence on every device, as under pmap); the mean is taken in device order."""

    def __init__(self, L):
        import threading
        self.L, self.slots, self.barrier, self.local = L, [None] * L, threading.Barrier(L), threading.local()

    def __call__(self, x, axis_name=None):
        self.slots[self.local.dev] = x
        self.barrier.wait()
        mean = self.slots[0]
        for y in self.slots[1:]:
            mean = mean + y
        mean = mean / self.L
        self.barrier.wait()
        return mean


def pmap_threads(ns, fn, per_device_args):
    """The reference's `jax.pmap(single_device_update, axis_name="local_devices")` (cleanba_ppo.py:656-660) on emulated devices.
    `fn`: the function object or its name in `ns`."""
    import threading
    L = len(per_device_args)
    pm = PMean(L)
    saved = ns["jax"]       # other threads (actors) keep using the namespace meanwhile: the swapped-in module is a superset
    ns["jax"] = NS(**{**saved.__dict__, "lax": NS(**{**saved.lax.__dict__, "pmean": pm})})
    results, errors = [None] * L, []
    f = ns[fn] if isinstance(fn, str) else fn

    def work(l):
        pm.local.dev = l
        torch.set_default_dtype(torch.float32)
        try:
            results[l] = f(*per_device_args[l])
        except BaseException as e:  # noqa: BLE001
            errors.append(e)
            pm.barrier.abort()
    threads = [threading.Thread(target=work, args=(l,)) for l in range(L)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    ns["jax"] = saved
    if errors:
        raise errors[0]
    return results


def cols(t, c):
    """Env-axis slice of a Transition (what prepare_data / device_put_sharded hand to one learner device)."""
    return type(t)(*[x[:, c] for x in t])


def run_ppo_update(ppo, rng):
    T, Bl, nmb, epochs = 4, 8, 4, 2
    a = ref_args(ppo, num_minibatches=nmb, update_epochs=epochs, num_steps=T, local_num_envs=Bl, num_actor_threads=1)
    a.num_updates = 1000
    ns = base_ns(args=a)
    for fn in ("linear_schedule", "get_logprob_entropy_value", "compute_gae_once"):
        run(lift(ppo, fn), ns)
    run(lift_stmt(ppo, lambda n: isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") == "compute_gae_once"), ns)
    for fn in ("compute_gae", "ppo_loss", "single_device_update"):
        run(lift(ppo, fn), ns)
    run(lift(ppo, "Transition"), ns)
    Transition = ns["Transition"]
    flat = net.init_params(11)
    halves = []
    fields = {}
    for h in range(2):      # two actor-thread payloads, hstack'ed by the update (cleanba_ppo.py:587)
        f = dict(obs=frames(600 + h, (T, Bl // 2, 4, 84, 84)), dones=rng.random((T, Bl // 2)) < 0.2,
                 actions=rng.integers(0, 18, (T, Bl // 2)).astype(np.int32), logprobs=(-2.9 + 0.2 * rng.standard_normal((T, Bl // 2))).astype(np.float32),
                 values=(0.3 * rng.standard_normal((T, Bl // 2))).astype(np.float32), rewards=rng.choice([-1.0, 0.0, 1.0], size=(T, Bl // 2)).astype(np.float32),
                 next_obs=frames(610 + h, (Bl // 2, 4, 84, 84)), next_done=rng.random(Bl // 2) < 0.3)
        fields[h] = f
        z = J(np.zeros((T, Bl // 2), np.float32))
        halves.append(Transition(obs=J(f["obs"]), dones=J(f["dones"]), actions=J(f["actions"]), logprobs=J(f["logprobs"]), values=J(f["values"]),
                                 env_ids=z, rewards=J(f["rewards"]), truncations=z, terminations=z, firststeps=z))
    key = tf.split(tf.PRNGKey(9), 4)[0]
    state = TrainState(J(flat), optim.Adam(flat.size), a.max_grad_norm, ns["linear_schedule"])
    state, loss, pg, vl, el, kl, key2 = ns["single_device_update"](state, halves, [J(fields[0]["next_obs"]), J(fields[1]["next_obs"])],
                                                                    [J(fields[0]["next_done"]), J(fields[1]["next_done"])], key)
    out = {f"upd_ppo_{k}{h}": v for h in range(2) for k, v in fields[h].items() if k not in ("obs", "next_obs")}
    out.update(upd_ppo_obs_seeds=np.array([600, 601], np.int64), upd_ppo_next_obs_seeds=np.array([610, 611], np.int64))
    out.update(upd_ppo_cfg=np.array([T, Bl, nmb, epochs, 1000], np.int64), upd_ppo_params_seed=np.int64(11), upd_ppo_key=key, upd_ppo_key_after=np.asarray(key2),
               upd_ppo_scalars=np.array([float(loss), float(pg), float(vl), float(el), float(kl)]), **digest("upd_ppo", flat, t2n(state.flat).astype(np.float32)),
               upd_ppo_opt_count=np.int64(state.opt.count))
    # gradient_accumulation_steps = 2 (cleanba_ppo.py:78, 492-500, 607): the shuffled batch is cut into num_minibatches * 2 mini-steps
    a.gradient_accumulation_steps = 2
    st_k = TrainState(J(flat), optim.Adam(flat.size), a.max_grad_norm, ns["linear_schedule"], k=2)
    st_k, loss, pg, vl, el, kl, key_k = ns["single_device_update"](st_k, halves, [J(fields[0]["next_obs"]), J(fields[1]["next_obs"])],
                                                                    [J(fields[0]["next_done"]), J(fields[1]["next_done"])], key)
    a.gradient_accumulation_steps = 1
    out.update(upd_ppok2_scalars=np.array([float(loss), float(pg), float(vl), float(el), float(kl)]), upd_ppok2_key_after=np.asarray(key_k),
               upd_ppok2_opt_count=np.int64(st_k.opt.count), **digest("upd_ppok2", flat, t2n(st_k.flat).astype(np.float32)))
    # the same payloads on TWO learner devices (multi_device_update, cleanba_ppo.py:656-660): device l holds env columns
    # [2l, 2l + 2) of each actor thread's payload, normalises its advantages locally, shuffles with the same key; gradients and
    # the reported scalars are pmean'ed
    per_dev = []
    for l in range(2):
        c = slice(2 * l, 2 * l + 2)
        per_dev.append((TrainState(J(flat), optim.Adam(flat.size), a.max_grad_norm, ns["linear_schedule"]), [cols(h, c) for h in halves],
                        [J(fields[h]["next_obs"][c]) for h in range(2)], [J(fields[h]["next_done"][c]) for h in range(2)], key))
    res = pmap_threads(ns, "single_device_update", per_dev)
    (st0, loss, pg, vl, el, kl, key2), st1 = res[0], res[1][0]
    assert torch.equal(st0.flat, st1.flat), "replicas diverged"
    out.update(upd_ppo2_scalars=np.array([float(loss), float(pg), float(vl), float(el), float(kl)]), upd_ppo2_key_after=np.asarray(key2),
               upd_ppo2_opt_count=np.int64(st0.opt.count), **digest("upd_ppo2", flat, t2n(st0.flat).astype(np.float32)))
    return out


def run_impala_update(imp, rng):
    T1, Bl, nmb = 5, 8, 4
    a = ref_args(imp, num_minibatches=nmb, num_steps=T1 - 1, local_num_envs=Bl, num_actor_threads=1)
    a.num_updates = 1000
    ns = base_ns(args=a)
    for fn in ("linear_schedule", "get_logits_and_value", "policy_gradient_loss", "entropy_loss_fn", "impala_loss", "single_device_update", "Transition"):
        run(lift(imp, fn), ns)
    Transition = ns["Transition"]
    flat = net.init_params(12)
    halves, fields = [], {}
    for h in range(2):
        f = dict(obs=frames(700 + h, (T1, Bl // 2, 4, 84, 84)), dones=rng.random((T1, Bl // 2)) < 0.2,
                 actions=rng.integers(0, 18, (T1, Bl // 2)).astype(np.int32), logitss=(0.3 * rng.standard_normal((T1, Bl // 2, 18))).astype(np.float32),
                 rewards=rng.choice([-1.0, 0.0, 1.0], size=(T1, Bl // 2)).astype(np.float32), firststeps=rng.random((T1, Bl // 2)) < 0.15)
        fields[h] = f
        kw = {k: J(v) for k, v in f.items()}
        z = J(np.zeros((T1, Bl // 2), np.float32))
        for extra in Transition._fields:
            kw.setdefault(extra, z)
        halves.append(Transition(**kw))
    key = tf.split(tf.PRNGKey(10), 4)[0]
    state = TrainState(J(flat), optim.RMSPropPyTorchStyle(flat.size, decay=0.99, eps=0.01), a.max_grad_norm, ns["linear_schedule"])
    state, loss, pg, vl, el, key2 = ns["single_device_update"](state, halves, key)
    out = {f"upd_imp_{k}{h}": v for h in range(2) for k, v in fields[h].items() if k != "obs"}
    out.update(upd_imp_obs_seeds=np.array([700, 701], np.int64))
    out.update(upd_imp_cfg=np.array([T1, Bl, nmb, 1000], np.int64), upd_imp_params_seed=np.int64(12),
               upd_imp_scalars=np.array([float(loss), float(pg), float(vl), float(el)]), **digest("upd_imp", flat, t2n(state.flat).astype(np.float32)),
               upd_imp_opt_count=np.int64(state.opt.count))
    # two learner devices (cleanba_impala.py:641-645): summed losses per device, pmean'ed gradients
    per_dev = []
    for l in range(2):
        c = slice(2 * l, 2 * l + 2)
        per_dev.append((TrainState(J(flat), optim.RMSPropPyTorchStyle(flat.size, decay=0.99, eps=0.01), a.max_grad_norm, ns["linear_schedule"]),
                        [cols(h, c) for h in halves], key))
    res = pmap_threads(ns, "single_device_update", per_dev)
    (st0, loss, pg, vl, el, _), st1 = res[0], res[1][0]
    assert torch.equal(st0.flat, st1.flat), "replicas diverged"
    out.update(upd_imp2_scalars=np.array([float(loss), float(pg), float(vl), float(el)]), upd_imp2_opt_count=np.int64(st0.opt.count),
               **digest("upd_imp2", flat, t2n(st0.flat).astype(np.float32)))
    return out


class Sharded(list):
    """What jax.device_put_sharded / device_put_replicated return here: one entry per (emulated) device."""


def _unshard(x, l):
    if isinstance(x, Sharded):
        return x[l]
    if isinstance(x, tuple) and hasattr(x, "_fields"):
        return type(x)(*[_unshard(e, l) for e in x])
    if isinstance(x, (list, tuple)):
        return type(x)(_unshard(e, l) for e in x)
    return x


def pmap_one_device(f, axis_name=None, devices=None):
    """jax.pmap over ONE learner device: strip the device axis from the arguments, put it back on the results."""
    def g(*args):
        out = f(*[_unshard(a, 0) for a in args])
        return tuple(o.reshape((1,) + tuple(o.shape)) if isinstance(o, torch.Tensor) else (Sharded([o]) if isinstance(o, np.ndarray) else o) for o in out)
    return g


class Rep:
    """flax.jax_utils.replicate(agent_state, devices): one train state per (emulated) learner device."""

    def __init__(self, states):
        self.states = list(states)

    params = property(lambda self: Sharded([st.params for st in self.states]))
    opt_state = property(lambda self: self.states[0].opt_state)
    flat = property(lambda self: self.states[0].flat)
    opt = property(lambda self: self.states[0].opt)


def pmap_devices(ns, f, devices):
    """jax.pmap over several emulated learner devices (threads + pmean rendezvous)."""
    L = len(devices)
    if L == 1:
        return pmap_one_device(f)

    def g(*args):
        per = [[a.states[l] if isinstance(a, Rep) else _unshard(a, l) for a in args] for l in range(L)]
        res = pmap_threads(ns, f, per)
        out = []
        for parts in zip(*res):
            if isinstance(parts[0], torch.Tensor):
                out.append(torch.stack(list(parts)))
            elif isinstance(parts[0], np.ndarray):
                out.append(Sharded(list(parts)))
            else:
                out.append(Rep(parts))
        return tuple(out)
    return g


class _LrView:
    """agent_state.opt_state[2][1].hyperparams["learning_rate"][-1].item() (cleanba_ppo.py:737-739)"""

    def __init__(self, lr):
        self.hyperparams = {"learning_rate": [torch.tensor(float(lr))]}

    def __getitem__(self, i):
        return self


def run_main(tree, algo, tag=None, extra=None):
    """The reference's whole `if __name__ == "__main__":` block (cleanba_ppo.py:409-801 / cleanba_impala.py:449-...), executed as
    written: size derivation, seeding, train-state creation, actor threads (the reference's own rollout(), real threads, real
    size-1 queues), the learner loop with its policy-version accounting and logging, for two updates on tests/golden/tiny_env.py."""
    import pprint as _pprint
    import queue
    import random
    import threading
    import time
    import uuid
    from collections import deque
    from types import SimpleNamespace
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import tiny_env
    N, T, threads_n, seed = 4, 3, 2, 5
    scalars, started, thread_errors = [], [], []
    lock = threading.Lock()

    class Writer:
        def __init__(self, *a, **k):
            pass

        def add_scalar(self, name, value, step):
            with lock:
                scalars.append((name, float(value), int(step), threading.current_thread() is threading.main_thread()))

        def add_text(self, *a, **k):
            pass

        def close(self):
            pass

    class Thread(threading.Thread):
        def __init__(self, target=None, args=()):
            def guarded(*a):
                try:
                    with torch.no_grad():
                        target(*a)
                except BaseException as e:  # noqa: BLE001
                    thread_errors.append(e)
            super().__init__(target=guarded, args=args, daemon=True)
            started.append(self)

    flat0 = net.init_params(seed)

    class Net(Network):
        def init(self, key, x):
            return Params(J(flat0)).network_params

        def tabulate(self, *a, **k):
            return ""

    class Act(Actor):
        init = Net.init
        tabulate = Net.tabulate

    class Cri(Critic):
        init = Net.init
        tabulate = Net.tabulate

    class TS(TrainState):
        @staticmethod
        def create(apply_fn=None, params=None, tx=None):
            _, chain, k = tx
            (_, max_norm), (_, opt_name, kw) = chain[1]
            lr = kw["learning_rate"]
            opt = optim.Adam(flat0.size, eps=kw["eps"]) if opt_name == "adam" else optim.RMSPropPyTorchStyle(flat0.size, decay=kw["decay"], eps=kw["eps"])
            return TS(J(flat0), opt, max_norm, lr if callable(lr) else (lambda count: lr), k)

        def apply_gradients(self, grads):
            self_lr = np.float32(self.schedule(self.opt.count))
            new = TrainState.apply_gradients(self, grads)
            out = TS(new.flat, new.opt, new.max_grad_norm, new.schedule, new.k, new.mini, new.acc)
            out.opt_state = _LrView(self_lr)
            return out

    optax = NS(clip_by_global_norm=lambda c: ("clip", c), adam="adam", chain=lambda *t: ("chain", t),
               inject_hyperparams=lambda opt: (lambda **kw: ("inner", opt, kw)), MultiSteps=lambda tx, every_k_schedule=1: ("multi", tx, every_k_schedule))
    import copy

    def replicate(x, devices=None):
        if devices is None or len(devices) == 1:
            return x
        return Rep([TS(x.flat.clone(), copy.deepcopy(x.opt), x.max_grad_norm, x.schedule, x.k) for _ in devices])
    flax = NS(core=NS(FrozenDict=dict), jax_utils=NS(replicate=replicate, unreplicate=lambda x: x[0] if isinstance(x, Sharded) else x), serialization=NS())
    jx = NS(**{**jax.__dict__, "process_count": lambda: 1, "process_index": lambda: 0, "local_devices": lambda: [0, 1], "devices": lambda: [0, 1],
               "device_put_sharded": lambda xs, devices=None: Sharded(xs), "device_put_replicated": lambda x, devices: Sharded([x for _ in devices]),
               "device_put": lambda x, device=None: x, "pmap": lambda f, axis_name=None, devices=None: pmap_devices(ns, f, devices), "distributed": NS()})
    ns = base_ns(jax=jx, optax=optax, flax=flax, make_env=tiny_env.make_env, time=time, deque=deque, queue=queue, random=random, uuid=uuid,
                 threading=NS(Thread=Thread), SimpleNamespace=SimpleNamespace, SummaryWriter=Writer, pprint=lambda *a, **k: None, print=lambda *a, **k: None,
                 Network=Net, Actor=Act, Critic=Cri, AgentParams=lambda n, a_, c: Params(J(flat0)), TrainState=TS, rmsprop_pytorch_style="rmsprop",
                 dataclass=dataclasses.dataclass, field=dataclasses.field, os=NS(path=NS(basename=lambda p_: "cleanba"), environ={}), __file__="cleanba.py",
                 __name__="__main__")
    run(lift(tree, "Args"), ns)
    # tame hyper-parameters: this run pins control flow, not optimisation -- with the defaults (16 Adam steps on 6-frame minibatches,
    # RMSProp's +-10 * lr first step) rounding differences of 1e-7 between two correct implementations grow to 1e-2 within one update
    over = dict(local_num_envs=N, num_steps=T, num_actor_threads=threads_n, seed=seed, log_frequency=1, total_timesteps=2 * N * T * threads_n,
                num_minibatches=2)
    over.update(dict(update_epochs=1) if algo == "ppo" else dict(learning_rate=2e-5))
    over.update(extra or {})
    algo = tag or algo

    def cli(cls):
        a = cls()
        for k, v in over.items():
            setattr(a, k, v)
        return a
    ns["tyro"] = NS(cli=cli)
    run(lift(tree, "Transition"), ns)
    run(lift(tree, "rollout"), ns)
    main = next(n for n in tree.body if isinstance(n, ast.If) and "__main__" in ast.unparse(n.test))
    exec(compile(ast.Module(body=main.body, type_ignores=[]), "<reference __main__>", "exec"), ns)
    for t in started:
        t.join(timeout=60)
    assert not thread_errors, thread_errors
    assert not any(t.is_alive() for t in started), "an actor thread did not finish"
    state = ns["agent_state"]
    if isinstance(state, Rep):
        assert all(torch.equal(st.flat, state.states[0].flat) for st in state.states), "learner replicas diverged"
        state = state.states[0]
    learner_scalars = [(n, v, s) for n, v, s, is_main in scalars if is_main]
    actor_scalars = [(n, v, s) for n, v, s, is_main in scalars if not is_main]
    keep = ("charts/learning_rate", "losses/value_loss", "losses/policy_loss", "losses/entropy", "losses/approx_kl", "losses/loss")
    out = {f"main_{algo}_cfg": np.array([N, T, threads_n, seed, int(ns["args"].num_updates)], np.int64),
           f"main_{algo}_overrides": np.array(json.dumps(over)),
           f"main_{algo}_learner_scalar_names": np.array(json.dumps([n for n, _, _ in learner_scalars])),
           f"main_{algo}_learner_scalars": np.array([[v, s] for n, v, s in learner_scalars if n in keep], np.float64),
           f"main_{algo}_actor_scalar_names": np.array(json.dumps([n for n, _, _ in actor_scalars])),
           f"main_{algo}_actor_returns": np.array([[v, s] for n, v, s in actor_scalars if n in ("charts/avg_episodic_return", "charts/avg_episodic_length")], np.float64),
           f"main_{algo}_versions": np.array([int(ns["learner_policy_version"]), int(ns["actor_policy_version"]), int(ns["update"]), int(ns["global_step"])], np.int64),
           f"main_{algo}_opt_count": np.int64(state.opt.count)}
    out.update(digest(f"main_{algo}", flat0, t2n(state.flat).astype(np.float32)))
    return out


def run_benchmark_tool():
    """cleanrl_utils/benchmark.py is plain Python: run its `__main__` as written (dry run + SLURM rendering of a template of ours with
    all of its placeholders) and keep what it prints and renders.  `distutils` (gone in Python 3.12) is the only stand-in."""
    import contextlib
    import io
    import runpy
    import tempfile
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import tiny_env
    tdir = tempfile.mkdtemp()
    tpath = os.path.join(tdir, "template.slurm")
    open(tpath, "w").write(tiny_env.SLURM_TEMPLATE)
    argv = ["benchmark.py", "--env-ids", "Breakout-v5", "Pong-v5", "--command", "python -m cleanba_b200.cleanba_ppo --local-num-envs 60",
            "--num-seeds", "3", "--start-seed", "4", "--workers", "0", "--auto-tag", "False",
            "--slurm-template-path", tpath, "--slurm-gpus-per-task", "4", "--slurm-total-cpus", "50",
            "--slurm-ntasks", "2", "--slurm-nodes", "2"]
    du = types.ModuleType("distutils")
    duu = types.ModuleType("distutils.util")
    duu.strtobool = lambda v: 1 if str(v).lower() in ("y", "yes", "t", "true", "on", "1") else 0
    du.util = duu
    saved_mods = {k: sys.modules.get(k) for k in ("distutils", "distutils.util")}
    sys.modules["distutils"], sys.modules["distutils.util"] = du, duu
    saved_argv, cwd = sys.argv, os.getcwd()
    buf = io.StringIO()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            sys.argv = argv
            with contextlib.redirect_stdout(buf):
                runpy.run_path(os.path.join(REF, "cleanrl_utils", "benchmark.py"), run_name="__main__")
            files = os.listdir("slurm")
            script = open(os.path.join("slurm", [f for f in files if f.endswith(".slurm")][0])).read()
        finally:
            os.chdir(cwd)
            sys.argv = saved_argv
            for k, v in saved_mods.items():
                if v is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = v
    lines = buf.getvalue().splitlines()
    i0 = lines.index("======= commands to run:") + 1
    commands = [l for l in lines[i0:] if l.startswith("python -m")]
    argv[argv.index("--slurm-template-path") + 1] = "<tiny_env.SLURM_TEMPLATE>"
    return {"bench_tool_argv": np.array(json.dumps(argv[1:])), "bench_tool_commands": np.array(json.dumps(commands)), "bench_tool_slurm": np.array(script)}


def run_rollout(tree, algo, rng):
    """The reference's whole rollout() thread function (cleanba_ppo.py:226-406 / cleanba_impala.py:268-447) for three updates on
    tests/golden/tiny_env.py, two learner devices: storage order, done / truncation / first-step flags, the IMPALA carried row,
    prepare_data's split of the env axis, global_step and policy-version accounting, episodic-return bookkeeping, scalar names."""
    import queue
    import time
    from collections import deque
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import tiny_env
    N, T, L, updates = 4, 3, 2, 3
    a = ref_args(tree, local_num_envs=N, num_steps=T, num_actor_threads=2, learner_device_ids=[0, 1], log_frequency=1, seed=3)
    a.num_updates, a.world_size = updates - 1, 1          # the loop runs range(1, num_updates + 2)
    scalars = []
    writer = NS(add_scalar=lambda name, value, step: scalars.append((name, float(value), int(step))))
    jx = NS(**{**jax.__dict__, "process_index": lambda: 0, "device_put_sharded": lambda xs, devices=None: Sharded(xs)})
    ns = base_ns(args=a, jax=jx, make_env=tiny_env.make_env, time=time, deque=deque, queue=queue, print=lambda *x, **k: None)
    run(lift(tree, "Transition"), ns)
    run(lift(tree, "rollout"), ns)
    flat = [net.init_params(21), net.init_params(22), net.init_params(23)]
    pq, rq = queue.Queue(), queue.Queue()
    for f in flat:
        pq.put(Params(J(f)))
    key = tf.split(tf.PRNGKey(3), 4)[0]
    with torch.no_grad():
        ns["rollout"](key, a, rq, pq, writer, [0, 1], 1, 0)          # device_thread_id = 1 (seed offset), actor device 0
    out = {f"ro_{algo}_cfg": np.array([N, T, L, updates], np.int64), f"ro_{algo}_param_seeds": np.array([21, 22, 23], np.int64), f"ro_{algo}_key": key}
    u = 0
    while not rq.empty():
        payload = rq.get()
        if algo == "ppo":
            gs, ver, upd, st, nobs, ndone, _, dtid = payload
            for l in range(L):
                out[f"ro_ppo_u{u}_l{l}_next_obs_sum"] = np.int64(np.asarray(nobs[l]).astype(np.int64).sum())
                out[f"ro_ppo_u{u}_l{l}_next_done"] = np.asarray(ndone[l])
        else:
            gs, ver, upd, st, _, dtid = payload
        out[f"ro_{algo}_u{u}_meta"] = np.array([gs, ver, upd, dtid], np.int64)
        for field in st._fields:
            for l in range(L):
                x = t2n(getattr(st, field)[l])
                if field == "obs":                       # frames: row-wise checksums instead of the bytes
                    x = x.reshape(x.shape[0], x.shape[1], -1).astype(np.int64).sum(-1)
                out[f"ro_{algo}_u{u}_l{l}_{field}"] = x
        u += 1
    assert u == updates
    out[f"ro_{algo}_scalar_names"] = np.array(json.dumps([n for n, _, _ in scalars]))
    keep = ("charts/avg_episodic_return", "charts/avg_episodic_length")
    out[f"ro_{algo}_scalars"] = np.array([[v, s] for n, v, s in scalars if n in keep], np.float64)
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_exec.npz")
    vec = build()
    np.savez_compressed(path, **vec)
    print("wrote", path, os.path.getsize(path), "bytes,", len(vec), "arrays")
