"""ORACLE package -- test infrastructure, NOT product code.

CPU restatement of the DTLR DINO-DETR hot path used only as a checker:
only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import it.
"""
