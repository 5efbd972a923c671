"""oracle/ -- CPU restatement of superMC's hot path + recipes to build the real reference.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package; supermc_b200/ never does.
"""
