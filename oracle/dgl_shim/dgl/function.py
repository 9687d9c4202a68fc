"""TEST INFRASTRUCTURE -- the three ``dgl.function`` builtins the reference uses
(``rigid_docking_model.py:204, 274-283``): ``u_sub_v``, ``copy_edge`` (deprecated name of
``copy_e``) and ``mean``."""
import operator


class BinaryMessage:
    def __init__(self, lhs, rhs, op, lhs_field, rhs_field, out):
        self.lhs, self.rhs, self.op = lhs, rhs, op
        self.lhs_field, self.rhs_field, self.out = lhs_field, rhs_field, out


class CopyEdge:
    def __init__(self, field, out):
        self.field, self.out = field, out


class Reduce:
    def __init__(self, kind, msg, out):
        self.kind, self.msg, self.out = kind, msg, out


def u_sub_v(lhs_field, rhs_field, out):
    """edge feature ``out`` = src[lhs_field] - dst[rhs_field]."""
    return BinaryMessage('u', 'v', operator.sub, lhs_field, rhs_field, out)


def copy_edge(edge, out):
    return CopyEdge(edge, out)


copy_e = copy_edge


def mean(msg, out):
    return Reduce('mean', msg, out)


def sum(msg, out):  # noqa: A001 - mirrors the DGL name
    return Reduce('sum', msg, out)
