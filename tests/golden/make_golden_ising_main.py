"""Runs the UNMODIFIED reference script main_MFQ_Ising.py over the UNMODIFIED reference environment (CPU, under the
gym 0.9 / imp stand-ins of oracle/ising_ref_shim) and keeps its standard output as a golden fixture:

    tests/golden/ising_main_n400_t0.8_ts50.txt     python main_MFQ_Ising.py -n 400 -t 0.8 -ts 50
    tests/golden/ising_main_n100_t0.25_ts40_ac0.6.txt   ... -n 100 -t 0.25 -ts 40 -ac 0.6 -dg 7

tests/test_ising_env.py replays the same command over the CUDA environment (python/examples/ising_model) and requires
the same lines: every random draw of the script comes from numpy's global generator (seed 13, main_MFQ_Ising.py:11),
so the two runs print identical text exactly when reset / step return identical observations and rewards.

    python tests/golden/make_golden_ising_main.py        (dev container only: needs /root/reference)
"""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
CASES = {"ising_main_n400_t0.8_ts50.txt": ["-n", "400", "-t", "0.8", "-ts", "50"],
         "ising_main_n100_t0.25_ts40_ac0.6.txt": ["-n", "100", "-t", "0.25", "-ts", "40", "-ac", "0.6", "-dg", "7"]}


def main():
    env = dict(os.environ, PYTHONPATH=os.path.join(REPO, "oracle", "ising_ref_shim"), PYTHONWARNINGS="ignore")
    for name, args in CASES.items():
        with tempfile.TemporaryDirectory() as cwd:     # the script creates ./ising_figs/<stamp>/display.npy
            out = subprocess.run([sys.executable, os.path.join(REF, "main_MFQ_Ising.py")] + args, cwd=cwd, env=env,
                                 stdout=subprocess.PIPE, check=True).stdout.decode()
        with open(os.path.join(HERE, name), "w") as f:
            f.write(out)
        print(name, len(out.splitlines()), "lines;", out.splitlines()[-1])


if __name__ == "__main__":
    main()
