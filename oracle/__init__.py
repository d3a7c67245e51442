"""ORACLE - test infrastructure only (CPU restatement of the reference's MPC QP path).

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` leg only.  Never imported by pympc_quadruped_b200.
"""
