"""Evaluation battles between two trained models -- the reference's battle.py with the same flags, over the CUDA
engine and the PyTorch learners (checkpoints written by train_battle.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

BASE_DIR = os.path.dirname(os.path.abspath(__file__))


def main(argv=None):
    from mfmarl_b200.cli import battle_parser, data_dirs
    args = battle_parser("evaluation battles between two trained models", training=False).parse_args(argv)
    args.data_dir, render_dir = data_dirs(args, BASE_DIR)

    import magent
    from mfmarl_b200.algo import spawn_ai, tools
    from mfmarl_b200.senario_battle import battle
    env = magent.GridWorld('battle', map_size=args.map_size)
    env.set_render_dir(render_dir)
    handles = env.get_handles()
    main_model_dir = os.path.join(args.data_dir, 'models/{}-0'.format(args.algo))
    oppo_model_dir = os.path.join(args.data_dir, 'models/{}-1'.format(args.oppo))
    models = [spawn_ai(args.algo, env, handles[0], args.algo + '-me', args.max_steps, device=args.device),
              spawn_ai(args.oppo, env, handles[1], args.oppo + '-opponent', args.max_steps, device=args.device)]
    models[0].load(main_model_dir, step=args.idx[0])
    models[1].load(oppo_model_dir, step=args.idx[1])
    runner = tools.Runner(env, handles, args.map_size, args.max_steps, models, battle, render_every=0)
    win_cnt = {'main': 0, 'opponent': 0}
    for k in range(0, args.n_round):
        runner.run(0.0, k, win_cnt=win_cnt)
    print('\n[*] >>> WIN_RATE: [{0}] {1} / [{2}] {3}'.format(args.algo, win_cnt['main'] / args.n_round, args.oppo,
                                                             win_cnt['opponent'] / args.n_round))
    return win_cnt


if __name__ == '__main__':
    main()
