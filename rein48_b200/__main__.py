# -*- coding: utf-8 -*-
"""`python -m rein48_b200 -c rand [-v y]`: the reference's CLI (main.py:51-79) over the GPU-backed
`Game` adapter -- one game with the random policy, final board printed.  Extra flags run the
batched form: `--episodes N` plays N fused rollouts and prints the statistics.

The keyboard policy (`-c hand`, control/hand.py) is interactive and not part of the data path;
it is accepted here for flag compatibility and reads moves from stdin like the reference."""
import argparse
import json
import random

from . import Game, EpisodeStats, play, random_rollouts


def hand_control(*_):
    """control/hand.py:7-21: ask until a valid spelling arrives."""
    from .game import action_code
    while True:
        print("Input action direction, then press ENTER button: ", end="")
        action = input()
        try:
            action_code(action)
            return action
        except ValueError:
            print("Input action signal is invalid, you must input valid value...")


def main(argv=None):
    ap = argparse.ArgumentParser(description="Play terminal 2048 on the GPU backend...")
    ap.add_argument("-c", "--control", type=str, dest="control", default="hand")
    ap.add_argument("-v", "--visual", type=str, dest="visual", default="y")
    ap.add_argument("--seed", type=int, default=None, help="random.seed() before the game (config 1)")
    ap.add_argument("--episodes", type=int, default=0, help="play this many fused random rollouts instead")
    ap.add_argument("--device", default="cuda:0")
    args = ap.parse_args(argv)
    control = "rand" if args.control in ("rand", "Rand", "RAND", "r", "R") else "hand"
    visual = args.visual in ("Y", "y", "Yes", "yes")
    if args.episodes:
        res = random_rollouts(args.episodes, seed=args.seed or 0, device=args.device)
        print(json.dumps(EpisodeStats(res.stats).summary()))
        return 0
    if args.seed is not None:
        random.seed(args.seed)
    game = Game(device=args.device)
    if control == "rand":
        score = play(game, "rand", show_result=visual)
    else:
        over = False
        while not over:
            Game.print_terminal(game.state_matrix)
            _, _, over = game.step(hand_control(game.state_matrix))
        Game.print_terminal(game.state_matrix)
        score = sum(sum(r) for r in game.state_matrix)
    print("score", int(score))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
