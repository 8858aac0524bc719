"""CPU oracles for the vessel-graph hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import anything from here; the product package (octa_autosegmentation_b200) never does.
"""
