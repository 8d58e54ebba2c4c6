"""TEST INFRASTRUCTURE (oracle) — not product code; never imported by gym_softrobot_b200.

NumPy restatement of the part of COOMM that gym-softrobot's muscle-driven octopus envs call
(`coomm.actuations.muscles`: LongitudinalMuscle, TransverseMuscle, ApplyMuscles), exposed under the import
name ``coomm`` so that the *unmodified* reference env code in `/root/reference/gym_softrobot/envs/octopus`
runs on top of it and produces golden vectors (oracle/gen_golden.py).

Third-party dependency: coomm 0.1.1, git rev d33fa034fe69b2cb6b4481d4297e4f82f60e981b of
github.com/hanson-hschang/COOMM (branch `refactor-numba-hotloops`), pinned at `/root/reference/uv.lock:172-179`.
It is not in `/root/reference` and cannot be fetched offline, so this is the published model (Chang, Halder,
Shih, Naughton, Gazzola, Mehta: "Energy-shaping control of a muscular octopus arm moving in three dimensions",
Proc. R. Soc. A 479, 2023; and the layout of the public COOMM package) as recalled — **PARITY UNPINNED**.
What is recalled rather than verified is listed in oracle/README.md ("COOMM restatement").
"""
