"""CPU parity oracle — TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
`--impl reference` legs of bench.py; never by anything under hdl-deflate_b200/.
"""
