"""ORACLE -- TEST INFRASTRUCTURE ONLY.

Nothing in the product package (gym_pybullet_drones_b200/) may import from here.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs use it, and only as the checker / the CPU baseline.

Two tiers:
  tier 1  ref_loader.py  -- the UNMODIFIED reference under /root/reference, made
          importable in this image by the stand-in modules in oracle/standins/.
          Exists only in the build container; used to generate tests/golden/*.npz.
  tier 2  dyn_oracle.py  -- batched float64 NumPy restatement of the same
          equations, pinned against tier-1 golden vectors; travels to the GPU box.
"""
