#!/usr/bin/env python
"""Demo driver with the reference's command line (cokwa/bitHTM example.py:20-32):
a noisy cyclic sequence is fed to an SP+TM network and the number of bursting,
correctly predicted and incorrectly predicted columns is printed per step.

    python example.py [--epochs 100] [--input_patterns 100] [--input_dim 1000]
                      [--input_density 0.2] [--input_noise_probability 0.05]
                      [--column_dim 2048] [--cell_dim 32] [--quiet] [--seed S]

Same stream of np.random draws as the reference: with the same --seed both print the
same numbers when the reference is given the deterministic inhibition rule
(tests/golden/make_golden.py shows how).  Needs a CUDA device.
"""

import argparse
import time

import numpy as np

from bithtm_b200 import HierarchicalTemporalMemory

if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--epochs", type=int, default=100)
    parser.add_argument("--input_patterns", type=int, default=100)
    parser.add_argument("--input_dim", type=int, default=1000)
    parser.add_argument("--input_density", type=float, default=0.2)
    parser.add_argument("--input_noise_probability", type=float, default=0.05)
    parser.add_argument("--column_dim", type=int, default=2048)
    parser.add_argument("--cell_dim", type=int, default=32)
    parser.add_argument("--seed", type=int, default=None, help="np.random.seed (the reference is unseeded)")
    parser.add_argument("--quiet", action="store_true", help="print a summary per epoch instead of per step")
    parser.add_argument("--readback", action="store_true",
                        help="compute the metrics on the host from cell_prediction, as the reference does")
    args = parser.parse_args()

    if args.seed is not None:
        np.random.seed(args.seed)
    inputs = np.random.rand(args.input_patterns, args.input_dim) < args.input_density  # example.py:34
    htm = HierarchicalTemporalMemory(args.input_dim, args.column_dim, args.cell_dim)

    w_e = len(str(max(args.epochs - 1, 1)))
    w_p = len(str(max(args.input_patterns - 1, 1)))
    w_c = len(str(max(args.column_dim - 1, 1)))
    w_a = len(str(max(htm.spatial_pooler.active_columns - 1, 1)))
    prev_column_prediction = np.zeros(args.column_dim, dtype=bool)  # example.py:50 (only with --readback)
    start_time = time.time()
    for epoch in range(args.epochs):
        totals = np.zeros(3, dtype=np.int64)
        for input_index, curr_input in enumerate(inputs):
            noisy_input = curr_input ^ (np.random.rand(args.input_dim) < args.input_noise_probability)  # :52
            sp_state, tm_state = htm.process(noisy_input)
            if args.readback:  # the reference's own expressions (example.py:50, 55-57): reads cell_prediction back
                burstings = tm_state.active_column_bursting.sum()
                corrects = prev_column_prediction[sp_state.active_column].sum()
                incorrects = prev_column_prediction.sum() - corrects
                prev_column_prediction = tm_state.cell_prediction.max(axis=1)
            else:  # the same three counts from the step summary (counted on the device)
                m = tm_state.column_metrics
                burstings, corrects, incorrects = m["bursting"], m["correct"], m["incorrect"]
            totals += (burstings, corrects, incorrects)
            if not args.quiet:
                print(f"epoch {epoch:{w_e}d}, pattern {input_index:{w_p}d}: bursting columns: {burstings:{w_a}d}, "
                      f"correct columns: {corrects:{w_a}d}, incorrect columns: {incorrects:{w_c}d}")
        if args.quiet:
            print(f"epoch {epoch:{w_e}d}: bursting {totals[0]}, correct {totals[1]}, incorrect {totals[2]}")
    print(f"{time.time() - start_time} seconds.")
