#!/usr/bin/env python
"""Command-line front end with the flag surface of the reference's sph_manager.jl (F/sph_manager.jl:13-63):

    --generate --run --EOS isothermal|polytropic --ic_type <name> --kwargs k=v,k=v
    --snapID 1 --snapInterval 10 --keepSnaps true --showPlots true

    python -m astrophysical_sph_b200.sph_manager --generate --EOS isothermal --ic_type boss_bodenheimer --kwargs N=100000
    python -m astrophysical_sph_b200.sph_manager --run --EOS isothermal --ic_type boss_bodenheimer --keepSnaps false

`--run` drives libsph_b200.so (one B200); `--generate` writes snapshots/<ic_type>/bin/1snap.csv.
Extra (not in the reference): --root <dir> (default "."), --maxSteps <n>, --device <ordinal>.
"""
from __future__ import annotations

import argparse
import sys


def _bool(v: str) -> bool:
    # ArgParse.jl parses Bool arguments from the literals `true` / `false`
    if v.lower() in ("true", "1"):
        return True
    if v.lower() in ("false", "0"):
        return False
    raise argparse.ArgumentTypeError(f"invalid Bool value: {v}")


def parse_kwargs(text: str) -> dict:
    """F/sph_manager.jl:75-98: `k=v,k=v` -> Bool / Int / Float64 / String."""
    out = {}
    if not text:
        return out
    for kv in text.split(","):
        k, v = kv.split("=")
        lo = v.lower()
        if lo == "true":
            out[k] = True
        elif lo == "false":
            out[k] = False
        else:
            try:
                out[k] = int(v)
            except ValueError:
                try:
                    out[k] = float(v)
                except ValueError:
                    out[k] = v
    return out


def parse_command_line(argv=None):
    p = argparse.ArgumentParser(prog="sph_manager", allow_abbrev=False)
    p.add_argument("--generate", action="store_true", help="Generate initial conditions only")
    p.add_argument("--run", action="store_true", help="Run simulation")
    p.add_argument("--EOS", type=str, required=True, help="Equation of State: isothermal or polytropic")
    p.add_argument("--ic_type", type=str, required=True,
                   help="Type of initial condition. Available options: sample_isothermal_sphere, sample_plummer_sphere, "
                        "bonnor_ebert_sphere, turbulent_molecular_cloud, rotating_cloud, polytropic_sphere, "
                        "gaussian_sphere, boss_bodenheimer")
    p.add_argument("--kwargs", type=str, default="",
                   help="Extra keyword arguments for initial conditions, in format key1=val1,key2=val2")
    p.add_argument("--snapID", type=int, default=1, help="Snapshot number to use for cold/warm start")
    p.add_argument("--snapInterval", type=int, default=10,
                   help="Interval in which we take a single snapshot of the simulation")
    p.add_argument("--keepSnaps", type=_bool, default=True, help="Keep or not the snapshots")
    p.add_argument("--showPlots", type=_bool, default=True, help="Only useful when keepSnaps is active")
    p.add_argument("--root", type=str, default=".", help="directory that holds snapshots/ (extension)")
    p.add_argument("--maxSteps", type=int, default=None, help="stop after this many loop iterations (extension)")
    p.add_argument("--device", type=int, default=0, help="CUDA device ordinal (extension)")
    return p.parse_args(argv)


def main(argv=None):
    args = parse_command_line(argv)
    if args.generate:
        from . import iniconds

        print(f"Generating {args.EOS} initial conditions for the test case of : {args.ic_type}")
        iniconds.iniconds_setup(args.EOS, args.ic_type, root=args.root, **parse_kwargs(args.kwargs))
    if args.run:
        if args.EOS in ("isothermal", "polytropic"):
            from . import sim

            print(f"Running {args.EOS} simulation from snapshot {args.snapID} with IC type: {args.ic_type}")
            sim.run_simulation(args.EOS, args.ic_type, args.snapID, args.snapInterval, args.keepSnaps, args.showPlots,
                               root=args.root, device=args.device, max_steps=args.maxSteps)
        else:
            print(f"No EOS of type {args.EOS} exists. Available options are either: 'isothermal' or 'polytropic'")
    return 0


if __name__ == "__main__":
    sys.exit(main())
