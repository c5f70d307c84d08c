"""Evaluation battles between two trained models -- the reference's battle.py with the same flags, over the CUDA
engine and the PyTorch learners (checkpoints written by train_battle.py)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

BASE_DIR = os.path.dirname(os.path.abspath(__file__))


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--algo', type=str, choices={'ac', 'mfac', 'mfq', 'il'}, required=True,
                        help='choose an algorithm from the preset')
    parser.add_argument('--oppo', type=str, choices={'ac', 'mfac', 'mfq', 'il'}, help='indicate the opponent model')
    parser.add_argument('--n_round', type=int, default=50, help='set the trainning round')
    parser.add_argument('--render', action='store_true', help='render or not (if true, will render every save)')
    parser.add_argument('--map_size', type=int, default=40, help='set the size of map')
    parser.add_argument('--max_steps', type=int, default=400, help='set the max steps')
    parser.add_argument('--idx', nargs='*', required=True)
    parser.add_argument('--device', type=str, default=None)
    parser.add_argument('--data_dir', type=str, default=os.path.join(BASE_DIR, 'data'))
    args = parser.parse_args(argv)

    import magent
    from mfmarl_b200.algo import spawn_ai, tools
    from mfmarl_b200.senario_battle import battle
    env = magent.GridWorld('battle', map_size=args.map_size)
    os.makedirs(os.path.join(args.data_dir, 'render'), exist_ok=True)
    env.set_render_dir(os.path.join(args.data_dir, 'render'))
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
