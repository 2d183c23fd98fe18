"""TEST INFRASTRUCTURE ONLY -- CPU restatements of the reference's Chambolle-Pock path.

Nothing under raocp-toolbox_b200/ imports this package.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it, and only as the checker / the reported CPU baseline.
"""
