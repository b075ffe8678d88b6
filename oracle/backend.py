"""CPU backend for the Sebulba plumbing (cleanba_b200.sebulba), built on the oracle.  TEST INFRASTRUCTURE: used by
tests/ (plumbing determinism, world_size-2 gloo) and by bench.py's CPU arm only; the product never imports it."""
import numpy as np

from . import impala as oimpala
from . import network as net
from . import ppo as oppo
from . import threefry as tf


class _Storage:
    def __init__(self, rows, N, impala):
        self.rows, self.impala = rows, impala
        self.obs = np.zeros((rows, N, 4, 84, 84), np.uint8)
        self.actions = np.zeros((rows, N), np.int32)
        self.logitss = np.zeros((rows, N, 18), np.float32)
        self.logprobs = np.zeros((rows, N), np.float32)
        self.values = np.zeros((rows, N), np.float32)
        self.host = {k: np.zeros((rows, N), dt) for k, dt in (("dones", bool), ("rewards", np.float32), ("firststeps", bool),
                                                              ("truncations", bool), ("terminations", np.int32), ("env_ids", np.int32))}

    def put_host(self, t, **fields):
        for k, v in fields.items():
            self.host[k][t] = v

    def take_carry(self):
        last = self.rows - 1
        return {"obs": self.obs[last].copy(), "actions": self.actions[last].copy(), "logitss": self.logitss[last].copy(),
                "host": {k: v[last].copy() for k, v in self.host.items()}}

    def put_carry(self, c):
        self.obs[0], self.actions[0], self.logitss[0] = c["obs"], c["actions"], c["logitss"]
        for k, v in c["host"].items():
            self.host[k][0] = v


class OracleActor:
    def __init__(self, N, args, key):
        self.N, self.impala, self.key, self.params = N, args.algo == "impala", np.asarray(key).copy(), None

    def new_storage(self, rows):
        return _Storage(rows, self.N, self.impala)

    def set_params(self, handle):
        self.params = handle

    def step(self, storage, t, obs_host):
        obs = np.asarray(obs_host)
        if self.impala:
            _, a, logits, self.key = oimpala.get_action(self.params, obs, self.key)
            storage.logitss[t] = logits
        else:
            _, a, lp, v, self.key, _ = oppo.get_action_and_value(self.params, obs, self.key)
            storage.logprobs[t], storage.values[t] = lp, v
        storage.obs[t], storage.actions[t] = obs, a
        return a, 0.0

    def shard_to_learners(self, storage, next_obs, next_done, L):
        N = self.N
        out = []
        for l in range(L):
            c = slice(l * N // L, (l + 1) * N // L)
            sh = {k: getattr(storage, k)[:, c] for k in ("obs", "actions", "logitss", "logprobs", "values")}
            sh.update({k: storage.host[k][:, c] for k in ("dones", "rewards", "firststeps")})
            if next_obs is not None:
                sh["next_obs"], sh["next_done"] = np.asarray(next_obs)[c], np.asarray(next_done)[c]
            out.append(sh)
        return out


class OracleLearner:
    def __init__(self, args, key, allreduce):
        self.args, self.impala = args, args.algo == "impala"
        params = net.init_params(args.seed, net.nature_param_spec() if getattr(args, "network", "impala_resnet") == "nature_cnn" else None)
        self.L = len(args.learner_device_ids)
        self.key = np.asarray(key).copy()
        self.allreduce = allreduce
        if self.impala:
            cfg = oimpala.ImpalaConfig(num_minibatches=args.num_minibatches, gamma=args.gamma, ent_coef=args.ent_coef,
                                       vf_coef=args.vf_coef, max_grad_norm=args.max_grad_norm, learning_rate=args.learning_rate,
                                       anneal_lr=args.anneal_lr, num_updates=max(args.num_updates, 1),
                                       gradient_accumulation_steps=getattr(args, "gradient_accumulation_steps", 1))
            self.learner = oimpala.ImpalaLearner(params, cfg)
        else:
            cfg = oppo.PPOConfig(num_minibatches=args.num_minibatches, update_epochs=args.update_epochs, gamma=args.gamma,
                                 gae_lambda=args.gae_lambda, clip_coef=args.clip_coef, ent_coef=args.ent_coef, vf_coef=args.vf_coef,
                                 max_grad_norm=args.max_grad_norm, learning_rate=args.learning_rate, anneal_lr=args.anneal_lr,
                                 norm_adv=args.norm_adv, num_updates=max(args.num_updates, 1),
                                 gradient_accumulation_steps=getattr(args, "gradient_accumulation_steps", 1))
            self.learner = oppo.PPOLearner(params, cfg)
        self.learner.cross_allreduce = allreduce
        self.world = max(args.world_size, 1)

    def update(self, payloads):
        shards = []
        for l in range(self.L):
            parts = [p[l] for p in payloads]
            cat = lambda k: np.concatenate([s[k] for s in parts], axis=1)
            if self.impala:
                shards.append(oimpala.Shard(obs=cat("obs"), dones=cat("dones"), actions=cat("actions"), logitss=cat("logitss"),
                                            rewards=cat("rewards"), firststeps=cat("firststeps")))
            else:
                shards.append(oppo.Shard(obs=cat("obs"), dones=cat("dones"), actions=cat("actions"), logprobs=cat("logprobs"),
                                         values=cat("values"), rewards=cat("rewards"),
                                         next_obs=np.concatenate([s["next_obs"] for s in parts]),
                                         next_done=np.concatenate([s["next_done"] for s in parts])))
        self.last_record = []
        if self.impala:
            return self.learner.update(shards, record=self.last_record)
        stats, self.key = self.learner.update(shards, self.key, record=self.last_record)
        return stats

    def params_for_actor(self, actor_device_id):
        return self.learner.params.copy()

    def stats_to_host(self, stats):
        names = ("loss", "pg_loss", "v_loss", "entropy_loss") + (() if self.impala else ("approx_kl",))
        return {k: float(v) for k, v in zip(names, stats)}

    def current_lr(self):
        """The rate the last optimizer step used (what the reference logs, cleanba_ppo.py:737-739)."""
        from . import optim
        a = self.args
        spu = a.num_minibatches * (1 if self.impala else a.update_epochs)
        return float(optim.linear_schedule(max(self.learner.opt.count - 1, 0), a.learning_rate, spu, max(a.num_updates, 1), a.anneal_lr))


class OracleBackend:
    def first_key(self, seed):
        return tf.split(tf.PRNGKey(seed), 4)[0]

    def make_learner(self, args, key, allreduce):
        return OracleLearner(args, key, allreduce)

    def make_actor(self, device_id, N, args, key):
        return OracleActor(N, args, key)
