"""CPU oracle for the CMax loss path - TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import anything from this package; the product (motionpriorcmax_b200) never does.
"""
