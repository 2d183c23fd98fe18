"""Node-type tags (API parity with reference raocp/core/nodes.py:3-31)."""


class Node:
    """Base tag: neither leaf nor nonleaf."""
    is_nonleaf = False
    is_leaf = False


class Nonleaf(Node):
    is_nonleaf = True


class Leaf(Node):
    is_leaf = True
