#!/usr/bin/env python
"""``python profiling.py --model VGG16 [--size B]`` → profiling.json (reference profiling.py:14-18)."""
import argparse

from split_learning_b200.profiler import write_profile, write_speed_profile

parser = argparse.ArgumentParser(description="Profiling Processing")
parser.add_argument("--model", type=str, required=True, help="Model name")
parser.add_argument("--size", type=int, required=False, default=4, help="Batch size")
parser.add_argument("--data", type=str, required=False, default=None)
parser.add_argument("--out", type=str, default="profiling.json")
parser.add_argument("--rounds", type=int, required=False, default=None,
                    help="FLEX-style profile (other/FLEX/profiling.py): whole-model speed over this many rounds only")
args = parser.parse_args()

if __name__ == "__main__" and args.rounds:
    info = write_speed_profile(args.model, args.size, args.out, args.data, args.rounds)
    print(f"End profiling: speed {info['speed']} samples/s")
elif __name__ == "__main__":
    info = write_profile(args.model, args.size, args.out, args.data)
    print(f"End profiling: {len(info['exe_time'])} layers, speed {info['speed']} samples/s, network {info['network']} B/ns")
