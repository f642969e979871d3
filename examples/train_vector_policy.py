"""Policy-gradient training on the GPU-resident vector env (agent integration example).

    python examples/train_vector_policy.py [--envs 16384] [--updates 30]

Bittner-28 target-control task (gym-PBN/Bittner-28-v0): each step the agent may flip one gene, the network then runs
until it reaches an attractor, reward +20 in the target attractor and -5 otherwise.  With `--multi` the env is the
reference's main one, the multi-action target env (gym-PBN/BittnerMulti-28-v0, README.md:14-17): up to three genes
flipped at once, chosen by three independent policy branches (the branching layout of the BDQ agent the reference is
used with), reward 1000 in the target attractor minus the number of interventions.  16 384 lockstep envs step in one
fused CUDA launch; observations, actions, rewards and the rollout buffer never leave the device.  A two-layer policy is
trained with advantage actor-critic on rollouts of 16 steps; the mean reward per step is printed per update.
"""
import argparse
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "gym-pbn-stac_b200"))

import gym_PBN  # noqa: E402
from gym_PBN.b200.rollout import RolloutBuffer  # noqa: E402
from gym_PBN.b200.vector_env import PBNVectorEnv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=16384)
    ap.add_argument("--updates", type=int, default=30)
    ap.add_argument("--horizon", type=int, default=16)
    ap.add_argument("--max-inner", type=int, default=256)
    ap.add_argument("--multi", action="store_true", help="multi-action target env with a three-branch policy")
    args = ap.parse_args()

    branches = 3 if args.multi else 1
    env = gym_PBN.make("gym-PBN/BittnerMulti-28-v0" if args.multi else "gym-PBN/Bittner-28-v0", seed=0,
                       max_inner_steps=args.max_inner)
    vec = PBNVectorEnv(env, args.envs, seed=1, action_slots=branches)
    n, n_actions = vec.n, vec.n + 1  # action 0 = no intervention, a = flip gene a-1 (pbn_target.py:257-262)
    dev = vec.device
    torch.manual_seed(0)
    body = torch.nn.Sequential(torch.nn.Linear(2 * n, 128), torch.nn.Tanh()).to(dev)
    pi_head, v_head = torch.nn.Linear(128, branches * n_actions).to(dev), torch.nn.Linear(128, 1).to(dev)
    params = list(body.parameters()) + list(pi_head.parameters()) + list(v_head.parameters())
    opt = torch.optim.Adam(params, lr=3e-3)
    buf = RolloutBuffer(vec, args.horizon, gamma=0.95, gae_lambda=0.9)

    def features(obs):
        # state bits plus the bits of the env's target state (info["target_state_packed"] of reset, refreshed on auto-reset)
        tgt = vec.sim.unpack(vec.sim.target_state)
        return torch.cat([obs.float(), tgt.float()], dim=-1) * 2 - 1

    def policy(obs):
        h = body(features(obs))
        dist = torch.distributions.Categorical(logits=pi_head(h).view(-1, branches, n_actions))  # one categorical per branch
        a = dist.sample()
        return a.to(torch.int32), dist.log_prob(a).sum(-1), v_head(h).squeeze(-1)

    obs, _ = vec.reset(seed=2)
    for u in range(args.updates):
        t0 = time.time()
        feats = []

        def recording_policy(o):
            feats.append(features(o))
            return policy(o)

        obs = buf.collect(recording_policy, obs)
        with torch.no_grad():
            last_v = v_head(body(features(obs))).squeeze(-1)
        adv, ret = buf.gae(last_v)
        x = torch.stack(feats)                                     # [T][B][2n]
        h = body(x)
        dist = torch.distributions.Categorical(logits=pi_head(h).view(x.shape[0], x.shape[1], branches, n_actions))
        logp = dist.log_prob(buf.actions[: buf.t].long()).sum(-1)
        v = v_head(h).squeeze(-1)
        a = (adv - adv.mean()) / (adv.std() + 1e-6)
        scale = 1000.0 if args.multi else 20.0  # reward scale of the env, keeps the value loss comparable
        loss = -(logp * a).mean() + 0.5 * ((v - ret) / scale).pow(2).mean() - 0.01 * dist.entropy().mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        torch.cuda.synchronize()
        dt = time.time() - t0
        print(f"update {u:3d}  mean reward/step {buf.rewards[: buf.t].mean().item():7.3f}  success rate "
              f"{(buf.rewards[: buf.t] > 0).float().mean().item():.3f}  {args.envs * args.horizon / dt:.3g} env-steps/s incl. learner")
    print("episode statistics:", vec.stats.summary() if hasattr(vec.stats, "summary") else vec.stats.v.tolist())


if __name__ == "__main__":
    main()
