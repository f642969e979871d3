"""The reference's example.py, on the B200 path: a PBCN from logic functions, a shipped Bittner network, a batched
vector env and an SSD estimate.  Run on a machine with a CUDA device after `python __graft_entry__.py`."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "gym-pbn-stac_b200"))

import gym_PBN  # noqa: E402
from gym_PBN.utils.eval import compute_ssd_hist  # noqa: E402


def iterate_through_env(env, action, steps):
    obs = env.reset()
    for _ in range(steps):
        obs, reward, terminated, truncated, info = env.step(action)
        print(obs, reward, terminated, truncated, info)
        if terminated or truncated:
            break


def example_1():
    """A custom PBCN from logic functions (example.py:19-44 of the reference; README-style goal_config)."""
    env = gym_PBN.make(
        "gym-PBN/PBCN-v0",
        logic_func_data=(
            ["u", "x1", "x2", "x3", "x4"],
            [[], [("not x2 and not x4", 1)], [("not x4 and not u and (x2 or x3)", 1)],
             [("not x2 and not x4 and x1", 0.7), ("False", 0.3)], [("not x2 and not x3", 1)]],
        ),
        goal_config={"all_attractors": [{(0, 0, 0, 0, 1)}, {(0, 0, 1, 0, 0)}], "target": {(0, 0, 0, 0, 1)}},
    )
    iterate_through_env(env, [1], 10)


def example_2():
    """A pre-packaged Bittner network (example.py:47-54)."""
    env = gym_PBN.make("gym-PBN/Bittner-200-v0")
    iterate_through_env(env, 0, 11)


def example_3():
    """65 536 lockstep copies of Bittner-100 on the GPU, then a steady-state-distribution estimate."""
    import torch

    vec = gym_PBN.make_vec("gym-PBN/Bittner-100-v0", num_envs=65536)
    obs, info = vec.reset(seed=0)
    for _ in range(10):
        actions = torch.randint(0, 101, (65536, 1), dtype=torch.int32, device=obs.device)
        obs, reward, terminated, truncated, info = vec.step(actions)
    print("episode statistics:", vec.stats.reduced())
    df, _ = compute_ssd_hist(vec.env, iters=1_200_000, resets=300)
    print(df.sort_values("Value", ascending=False).head())


if __name__ == "__main__":
    example_1()
    example_2()
    example_3()
