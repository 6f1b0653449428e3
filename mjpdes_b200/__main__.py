"""``python -m mjpdes_b200 -i <model.clj> [-o <out>]``: the reference CLI (main.clj:7-18) on the CUDA path."""
import argparse
import sys

from . import model


def main(argv=None):
    ap = argparse.ArgumentParser(prog="mjpdes_b200")
    ap.add_argument("-i", "--input-file", required=True)                 # main.clj:8-9
    ap.add_argument("-o", "--output-file")                               # main.clj:10
    ap.add_argument("--seed", type=lambda s: int(s, 0), default=model.DEFAULT_SEED)
    a = ap.parse_args(argv)
    m = model.read_model(a.input_file)
    out = open(a.output_file, "w") if a.output_file else sys.stdout
    try:
        model.main_loop(m, out_stream=out, seed=a.seed)
    finally:
        if a.output_file:
            out.close()


if __name__ == "__main__":
    main()
