#!/usr/bin/env python
"""``python profiling.py --model VGG16 [--size B]`` → profiling.json (reference profiling.py:14-18)."""
import argparse

from split_learning_b200.profiler import write_profile

parser = argparse.ArgumentParser(description="Profiling Processing")
parser.add_argument("--model", type=str, required=True, help="Model name")
parser.add_argument("--size", type=int, required=False, default=4, help="Batch size")
parser.add_argument("--data", type=str, required=False, default=None)
parser.add_argument("--out", type=str, default="profiling.json")
args = parser.parse_args()

if __name__ == "__main__":
    info = write_profile(args.model, args.size, args.out, args.data)
    print(f"End profiling: {len(info['exe_time'])} layers, speed {info['speed']} samples/s, network {info['network']} B/ns")
