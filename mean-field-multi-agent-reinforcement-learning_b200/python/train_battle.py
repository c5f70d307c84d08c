"""Self-play training on the battle scenario -- the reference's train_battle.py with the same flags, over the CUDA
engine and the PyTorch learners.

    python train_battle.py --algo mfq                   one environment through the magent binding (reference loop)
    python train_battle.py --algo mfq --envs 1024       1024 lock-stepped environments on the GPU per round:
                                                         observations, mean actions and replay stay in HBM
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

BASE_DIR = os.path.dirname(os.path.abspath(__file__))


def linear_decay(epoch, x, y):
    """piecewise-linear schedule through the points (x[i], y[i]) (train_battle.py:19-40)"""
    if epoch == x[0]:
        return y[0]
    eps = y[0]
    for i, x_i in enumerate(x):
        if epoch <= x_i:
            slope = (y[i] - y[i - 1]) / (x_i - x[i - 1])
            eps = slope * (epoch - x[i - 1]) + y[i - 1]
            break
    return eps


class BatchedEnvAdapter:
    """What Runner / spawn_ai ask of `env` (space queries) on top of a BatchedGridWorld, plus a `play` handle that
    reduces the per-environment statistics of play_batched to the scalars Runner logs (means over environments)."""

    def __init__(self, benv):
        self.benv = benv

    def get_view_space(self, handle):
        s = self.benv.sizes
        return (s["view_size"], s["view_size"], s["n_channel"])

    def get_feature_space(self, handle):
        return (self.benv.sizes["feature_size"],)

    def get_action_space(self, handle):
        return (self.benv.sizes["n_action"],)

    def play(self, env, n_round, map_size, max_steps, handles, models, print_every, eps=1.0, render=False, train=False):
        from mfmarl_b200.senario_battle import play_batched
        print("\n\n[*] ROUND #{0}, EPS: {1:.2f} ENVS: {2}".format(n_round, eps, self.benv.n_envs))
        max_nums, nums, mean_r, total_r = play_batched(self.benv, n_round, max_steps, models, eps=eps, train=train,
                                                       print_every=print_every)
        return (list(max_nums.mean(axis=0)), list(nums.mean(axis=0)), list(mean_r.mean(axis=0)),
                list(total_r.mean(axis=0)))


def main(argv=None):
    from mfmarl_b200.cli import battle_parser, data_dirs
    args = battle_parser("self-play training on the battle scenario", training=True).parse_args(argv)
    args.data_dir, render_dir = data_dirs(args, BASE_DIR)

    from mfmarl_b200.algo import spawn_ai, tools
    log_dir = os.path.join(args.data_dir, 'tmp')
    model_dir = os.path.join(args.data_dir, 'models/{}'.format(args.algo))
    if args.envs > 0:
        from mfmarl_b200 import BatchedGridWorld
        cap = max(64, int(args.map_size * args.map_size * 0.04))
        benv = BatchedGridWorld(args.envs, map_size=args.map_size, capacity=cap, device=args.device, rng="philox")
        env = BatchedEnvAdapter(benv)
        handles, play = [0, 1], env.play
    else:
        import magent
        from mfmarl_b200.senario_battle import play
        env = magent.GridWorld('battle', map_size=args.map_size)
        env.set_render_dir(render_dir)
        handles = env.get_handles()
    # batched rounds produce envs x agents x steps rows: give the device replay room for up to 2^20 of them (5 GB)
    device_rows = min(1 << 20, args.envs * cap * args.max_steps) if args.envs > 0 else None
    # ... and scale the Q learners' minibatch with the number of recorded environments, so that a round keeps the
    # reference's number of gradient steps (new rows * 2 / batch, tools.py:352-360) instead of multiplying it
    rec_envs = max(1, min(args.envs, device_rows // (cap * args.max_steps))) if args.envs > 0 else 1
    models = [spawn_ai(args.algo, env, handles[0], args.algo + '-me', args.max_steps, device=args.device,
                       device_rows=device_rows, batch_size=64 * rec_envs),
              spawn_ai(args.algo, env, handles[1], args.algo + '-opponent', args.max_steps, device=args.device)]
    runner = tools.Runner(env, handles, args.map_size, args.max_steps, models, play,
                          render_every=args.save_every if args.render else 0, save_every=args.save_every, tau=0.01,
                          log_name=args.algo, log_dir=log_dir, model_dir=model_dir, train=True)
    for k in range(0, args.n_round):
        eps = linear_decay(k, [0, int(args.n_round * 0.8), args.n_round], [1, 0.2, 0.1])
        runner.run(eps, k)
    return runner


if __name__ == '__main__':
    main()
