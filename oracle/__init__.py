"""CPU oracle package -- TEST INFRASTRUCTURE.  Importable only from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  See grasp_ik_np.py (numpy) and grasp_ik_oracle.c (C)."""
