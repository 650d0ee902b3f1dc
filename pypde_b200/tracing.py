"""Lowering of reference-style user functions to CUDA source by symbolic tracing.

The reference compiles `F(Q[, dQ], d) -> ndarray`, `B(Q[, d]) -> (V, V) ndarray`
and `S(Q) -> ndarray` with numba for the CPU (pypde/cfuncs.py:30-75).  numba's
CUDA target cannot allocate or return arrays, so such functions cannot be
lowered for the GPU as they stand.  They are,
however, straight-line arithmetic on a handful of scalars, so this module runs
them ONCE per direction on symbolic inputs — numpy object arrays of `Sym`
scalars, with the numpy / math names in the function's globals replaced by
tracing versions — and prints the recorded expression DAG as an
`extern "C" __device__` function, in the evaluation order Python used (so
+ - * / sqrt give the same bits as the CPU evaluation of the same function).

Supported inside user functions: indexing / slicing / arithmetic on arrays,
`zeros`, `ones`, `array`, `eye`, `dot`, `inner`, `outer`, `sum`, `trace`, `det`
(up to 3 x 3), `sqrt`, `exp`, `log`, `abs`, `sin`, `cos`, `tanh`, `maximum`, `minimum`,
`where`, `**` (small integer powers become products, as numba emits them),
helper functions (plain or `@njit`) called with array or scalar arguments,
branches on the direction `d` or on constants, and branches on *data* values
(`K0 if T > Ti else 0`, `if rho > rho0: return a` ... `return b`, if / else blocks
of assignments, `and` / `or` / `not` of comparisons): those are rewritten on the
source before tracing into both-arms-then-select form (see _lower_branches).  numpy
names must be module-level imports of the function's module (imports inside the
function body bind the real numpy and are not seen by the tracer).
"""
import math
import types

import numpy as np


class TraceError(TypeError):
    pass


_BIN = {'add': '+', 'sub': '-', 'mul': '*', 'div': '/'}
_FUN1 = ('sqrt', 'exp', 'log', 'fabs', 'sin', 'cos', 'tanh', 'neg')
_CMP = {'lt': '<', 'le': '<=', 'gt': '>', 'ge': '>=', 'eq': '==', 'ne': '!='}


class Tape:
    def __init__(self):
        self.nodes = []     # (op, args)
        self.cse = {}

    def node(self, op, *args):
        key = (op, ) + tuple(a.idx if isinstance(a, Sym) else ('c', float(a).hex())
                             if not isinstance(a, str) else a for a in args)
        if key in self.cse:
            return self.cse[key]
        s = Sym(self, len(self.nodes))
        self.nodes.append((op, args))
        self.cse[key] = s
        return s


def _num(x):
    return isinstance(x, (int, float, np.integer, np.floating)) and not isinstance(x, bool)


class Sym:
    """A scalar of the traced computation."""
    __slots__ = ('tape', 'idx')

    def __init__(self, tape, idx):
        self.tape, self.idx = tape, idx

    def _bin(self, op, a, b):
        if not (isinstance(a, Sym) or _num(a)) or not (isinstance(b, Sym) or _num(b)):
            return NotImplemented
        return self.tape.node(op, a if isinstance(a, Sym) else float(a),
                              b if isinstance(b, Sym) else float(b))

    def __add__(self, o):
        return self._bin('add', self, o)

    def __radd__(self, o):
        return self._bin('add', o, self)

    def __sub__(self, o):
        return self._bin('sub', self, o)

    def __rsub__(self, o):
        return self._bin('sub', o, self)

    def __mul__(self, o):
        return self._bin('mul', self, o)

    def __rmul__(self, o):
        return self._bin('mul', o, self)

    def __truediv__(self, o):
        return self._bin('div', self, o)

    def __rtruediv__(self, o):
        return self._bin('div', o, self)

    def __neg__(self):
        return self.tape.node('neg', self)

    def __pos__(self):
        return self

    def __abs__(self):
        return self.tape.node('fabs', self)

    def __pow__(self, e):
        if _num(e) and float(e) == int(e) and 0 <= int(e) <= 4:
            e = int(e)            # numba emits small integer powers as products
            if e == 0:
                return 1.0
            r = self
            for _ in range(e - 1):
                r = r * self
            return r
        if _num(e) and float(e) == 0.5:
            return self.tape.node('sqrt', self)
        return self.tape.node('pow', self, e if isinstance(e, Sym) else float(e))

    def __rpow__(self, b):
        return self.tape.node('pow', float(b), self)

    def _cmp(self, op, o):
        return SymBool(self.tape.node(op, self, o if isinstance(o, Sym) else float(o)))

    def __lt__(self, o):
        return self._cmp('lt', o)

    def __le__(self, o):
        return self._cmp('le', o)

    def __gt__(self, o):
        return self._cmp('gt', o)

    def __ge__(self, o):
        return self._cmp('ge', o)

    def __eq__(self, o):
        return self._cmp('eq', o)

    def __ne__(self, o):
        return self._cmp('ne', o)

    def __bool__(self):
        raise TraceError('a branch depends on a computed value; use where(cond, a, b)')

    def __float__(self):
        raise TraceError('a traced value was converted to a Python float (assignment into a '
                         'float array? create arrays with zeros/array/eye from numpy)')

    __hash__ = object.__hash__


class SymBool:
    __slots__ = ('node', )

    def __init__(self, node):
        self.node = node

    def __bool__(self):
        raise TraceError('a branch depends on a computed value in a place the tracer cannot '
                         'lower (inside a loop, or the function\'s source is unavailable); '
                         'write it as where(cond, a, b)')

    def _logic(self, op, o):
        if isinstance(o, SymBool):
            return SymBool(self.node.tape.node(op, self.node, o.node))
        if op == 'and':
            return self if o else False
        return True if o else self

    def __and__(self, o):
        return self._logic('and', o)

    __rand__ = __and__

    def __or__(self, o):
        return self._logic('or', o)

    __ror__ = __or__

    def __invert__(self):
        return SymBool(self.node.tape.node('not', self.node))


def _map(fn, x):
    if isinstance(x, np.ndarray):
        out = np.empty(x.shape, dtype=object)
        for i, v in np.ndenumerate(x):
            out[i] = fn(v)
        return out
    return fn(x)


def _fun1(name, pyfn):
    def f(x):
        def one(v):
            if isinstance(v, Sym):
                return v.tape.node(name, v)
            return pyfn(float(v))
        return _map(one, x)
    f.__name__ = name
    return f


def _obj(a):
    if isinstance(a, np.ndarray) and a.dtype == object:
        return a
    out = np.empty(np.shape(a), dtype=object)
    for i, v in np.ndenumerate(np.asarray(a, dtype=object)):
        out[i] = v if isinstance(v, Sym) else float(v)
    return out


def t_zeros(shape, dtype=None):
    a = np.empty(shape, dtype=object)
    a.fill(0.0)
    return a


def t_ones(shape, dtype=None):
    a = np.empty(shape, dtype=object)
    a.fill(1.0)
    return a


def t_array(x, dtype=None):
    return _obj(x).copy()


def t_eye(n, dtype=None):
    a = t_zeros((n, n))
    for i in range(n):
        a[i, i] = 1.0
    return a


def t_dot(a, b):
    a, b = _obj(a), _obj(b)
    if a.ndim == 1 and b.ndim == 1:
        acc = a[0] * b[0]
        for k in range(1, a.shape[0]):
            acc = acc + a[k] * b[k]
        return acc
    if a.ndim == 2 and b.ndim == 1:
        return np.array([t_dot(a[i], b) for i in range(a.shape[0])], dtype=object)
    if a.ndim == 1 and b.ndim == 2:
        return np.array([t_dot(a, b[:, j]) for j in range(b.shape[1])], dtype=object)
    out = np.empty((a.shape[0], b.shape[1]), dtype=object)
    for i in range(a.shape[0]):
        for j in range(b.shape[1]):
            out[i, j] = t_dot(a[i], b[:, j])
    return out


def t_outer(a, b):
    a, b = _obj(a).ravel(), _obj(b).ravel()
    out = np.empty((a.size, b.size), dtype=object)
    for i in range(a.size):
        for j in range(b.size):
            out[i, j] = a[i] * b[j]
    return out


def t_sum(a, axis=None):
    a = _obj(a)
    if axis is not None:
        return np.apply_along_axis(lambda r: t_sum(r), axis, a)
    flat = a.ravel()
    acc = flat[0]
    for v in flat[1:]:
        acc = acc + v
    return acc


def t_trace(a):
    a = _obj(a)
    acc = a[0, 0]
    for i in range(1, a.shape[0]):
        acc = acc + a[i, i]
    return acc


def t_det(a):
    a = _obj(a)
    n = a.shape[0]
    if n == 1:
        return a[0, 0]
    if n == 2:
        return a[0, 0] * a[1, 1] - a[0, 1] * a[1, 0]
    if n == 3:
        return (a[0, 0] * (a[1, 1] * a[2, 2] - a[1, 2] * a[2, 1]) -
                a[0, 1] * (a[1, 0] * a[2, 2] - a[1, 2] * a[2, 0]) +
                a[0, 2] * (a[1, 0] * a[2, 1] - a[1, 1] * a[2, 0]))
    raise TraceError('det of matrices larger than 3 x 3 is not supported by the tracer')


def _select(c, a, b):
    if isinstance(c, SymBool):
        tape = c.node.tape
        return tape.node('select', c.node, a if isinstance(a, Sym) else float(a),
                         b if isinstance(b, Sym) else float(b))
    return a if c else b


def t_where(c, a, b):
    if isinstance(c, np.ndarray) or isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        c, a, b = np.broadcast_arrays(np.asarray(c, dtype=object), _obj(a), _obj(b))
        out = np.empty(c.shape, dtype=object)
        for i in np.ndindex(c.shape):
            out[i] = _select(c[i], a[i], b[i])
        return out
    return _select(c, a, b)


def _minmax(op):
    def f(a, b):
        def one(x, y):
            if isinstance(x, Sym) or isinstance(y, Sym):
                tape = (x if isinstance(x, Sym) else y).tape
                return tape.node(op, x if isinstance(x, Sym) else float(x),
                                 y if isinstance(y, Sym) else float(y))
            return max(x, y) if op == 'fmax' else min(x, y)
        if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
            a, b = np.broadcast_arrays(_obj(a), _obj(b))
            out = np.empty(a.shape, dtype=object)
            for i in np.ndindex(a.shape):
                out[i] = one(a[i], b[i])
            return out
        return one(a, b)
    return f


TRACED = {
    'zeros': t_zeros, 'ones': t_ones, 'empty': t_zeros, 'array': t_array, 'asarray': t_array,
    'eye': t_eye, 'identity': t_eye, 'dot': t_dot, 'inner': t_dot, 'matmul': t_dot,
    'outer': t_outer, 'sum': t_sum, 'trace': t_trace, 'det': t_det, 'where': t_where,
    'sqrt': _fun1('sqrt', math.sqrt), 'exp': _fun1('exp', math.exp), 'log': _fun1('log', math.log),
    'abs': _fun1('fabs', abs), 'fabs': _fun1('fabs', abs), 'absolute': _fun1('fabs', abs),
    'sin': _fun1('sin', math.sin), 'cos': _fun1('cos', math.cos),
    'tanh': _fun1('tanh', math.tanh), 'maximum': _minmax('fmax'), 'minimum': _minmax('fmin'),
    'fmax': _minmax('fmax'), 'fmin': _minmax('fmin'),
}


class _Proxy:
    """Stands in for the numpy / math / numpy.linalg modules inside traced code."""

    def __init__(self, real):
        self._real = real

    def __getattr__(self, name):
        if name in TRACED:
            return TRACED[name]
        v = getattr(self._real, name)
        if isinstance(v, types.ModuleType):
            return _Proxy(v)
        if callable(v) and not isinstance(v, type):
            raise TraceError('%s.%s is not supported inside GPU user functions' %
                             (self._real.__name__, name))
        return v


class _UserModuleProxy:
    """A user's own module referenced from traced code (e.g. `mg.pressure(...)`):
    functions fetched through it are retargeted like directly imported ones."""

    def __init__(self, real, seen):
        self._real, self._seen = real, seen

    def __getattr__(self, name):
        v = getattr(self._real, name)
        if id(v) in _NP_IDS:
            return _NP_IDS[id(v)]
        if isinstance(v, types.ModuleType):
            if v.__name__.split('.')[0] in ('numpy', 'math'):
                return _Proxy(v)
            return _UserModuleProxy(v, self._seen)
        if hasattr(v, 'py_func') or isinstance(v, types.FunctionType):
            return retarget(v, self._seen)
        return v


_NP_IDS = {}
for _n, _f in TRACED.items():
    for _mod in (np, np.linalg, math):
        if hasattr(_mod, _n):
            _NP_IDS[id(getattr(_mod, _n))] = _f
_NP_IDS[id(abs)] = TRACED['abs']
_NP_IDS[id(sum)] = t_sum
_NP_IDS[id(max)] = TRACED['maximum']
_NP_IDS[id(min)] = TRACED['minimum']


# ---------------------------------------------------------------------------
# Data-dependent branches.  Python's `a if c else b`, `if c:`, `and`, `or`, `not` call
# bool(c), which a traced comparison cannot answer; the reference's own reactive Euler
# source does exactly that (pypde/tests/reactive_euler/system.py:47 `K0 if T > Ti else 0`).
# Before a function is traced its source is therefore rewritten (AST -> AST):
#   a if c else b          ->  __pde_ifexp(c, lambda: a, lambda: b)
#   x and y / x or y / not ->  __pde_and / __pde_or / __pde_not (lazy, like Python)
#   if c: A  else: B       ->  both branches run on copies of the variables they assign,
#                              then every such variable becomes where(c, its A value, its B value)
#   if c: ...return r      ->  the rest of the function body is moved into both arms:
#      rest...                 return __pde_ifexp(c, then_arm, else_arm)
# With an ordinary bool for c (branches on `d`, on constants) the helpers do what Python
# does, evaluating one arm only; with a traced condition both arms are evaluated and
# merged by select — the branch-free form GPU code wants anyway.  An `if` inside a loop,
# `with` or `try` is left alone (and reports the TraceError above if it is data dependent).
# ---------------------------------------------------------------------------
import ast
import copy as _copy
import inspect
import textwrap


class _Undef:
    def __repr__(self):
        return '<variable not assigned before a data-dependent if>'


_UNDEF = _Undef()


def _cp(v):
    return v.copy() if isinstance(v, np.ndarray) else v


def _pde_try(thunk):
    try:
        return _cp(thunk())
    except NameError:
        return _UNDEF


def _pde_plain(c):
    return not isinstance(c, SymBool)


def _pde_ifexp(c, fa, fb):
    if isinstance(c, SymBool):
        a, b = fa(), fb()
        if a is None or b is None:
            raise TraceError('one arm of a data-dependent branch returns nothing')
        return t_where(c, a, b)
    return fa() if c else fb()


def _pde_and(*thunks):
    acc = True
    for t in thunks:
        v = t()
        if isinstance(v, SymBool) or isinstance(acc, SymBool):
            acc = (acc & v) if isinstance(acc, SymBool) else (v & acc)
        else:
            acc = v
            if not v:
                return v
    return acc


def _pde_or(*thunks):
    acc = False
    for t in thunks:
        v = t()
        if isinstance(v, SymBool) or isinstance(acc, SymBool):
            acc = (acc | v) if isinstance(acc, SymBool) else (v | acc)
        else:
            acc = v
            if v:
                return v
    return acc


def _pde_not(v):
    return ~v if isinstance(v, SymBool) else (not v)


def _pde_restore(vals):
    return tuple(_cp(v) for v in vals)


def _pde_merge(c, a_vals, b_vals, names):
    out = []
    for a, b, n in zip(a_vals, b_vals, names):
        if a is _UNDEF or b is _UNDEF:
            raise TraceError('variable %r is assigned in only one arm of a data-dependent if '
                             'and not before it' % n)
        out.append(t_where(c, a, b))
    return tuple(out)


_PDE_HELPERS = {'__pde_try': _pde_try, '__pde_plain': _pde_plain, '__pde_ifexp': _pde_ifexp,
                '__pde_and': _pde_and, '__pde_or': _pde_or, '__pde_not': _pde_not,
                '__pde_restore': _pde_restore, '__pde_merge': _pde_merge, '__pde_cp': _cp}


def _stored_names(stmts):
    """Names bound or mutated (x = ..., x += ..., x[i] = ..., for x in ...) in a block, not
    descending into nested function definitions."""
    names = []

    def target(t):
        if isinstance(t, ast.Name):
            if t.id not in names:
                names.append(t.id)
        elif isinstance(t, (ast.Tuple, ast.List)):
            for e in t.elts:
                target(e)
        elif isinstance(t, (ast.Subscript, ast.Attribute, ast.Starred)):
            target(t.value)

    def visit(n):
        if isinstance(n, (ast.FunctionDef, ast.Lambda, ast.ClassDef)):
            return
        if isinstance(n, ast.Assign):
            for t in n.targets:
                target(t)
        elif isinstance(n, (ast.AugAssign, ast.AnnAssign, ast.For)):
            target(n.target)
        elif isinstance(n, ast.NamedExpr):
            target(n.target)
        for ch in ast.iter_child_nodes(n):
            visit(ch)

    for st in stmts:
        visit(st)
    return names


def _has_return(stmts):
    def visit(n):
        if isinstance(n, (ast.FunctionDef, ast.Lambda, ast.ClassDef)):
            return False
        if isinstance(n, ast.Return):
            return True
        return any(visit(ch) for ch in ast.iter_child_nodes(n))
    return any(visit(st) for st in stmts)


def _ends_in_return(stmts):
    if not stmts:
        return False
    last = stmts[-1]
    if isinstance(last, ast.Return):
        return True
    if isinstance(last, ast.If):
        return _ends_in_return(last.body) and _ends_in_return(last.orelse)
    return False


class _ExprRewriter(ast.NodeTransformer):
    """IfExp / BoolOp / Not -> helper calls with lazily evaluated arms."""

    @staticmethod
    def _thunk(e):
        return ast.Lambda(args=ast.arguments(posonlyargs=[], args=[], kwonlyargs=[],
                                             kw_defaults=[], defaults=[]), body=e)

    @staticmethod
    def _call(name, args):
        return ast.Call(func=ast.Name(id=name, ctx=ast.Load()), args=args, keywords=[])

    def visit_IfExp(self, n):
        self.generic_visit(n)
        return self._call('__pde_ifexp', [n.test, self._thunk(n.body), self._thunk(n.orelse)])

    def visit_BoolOp(self, n):
        self.generic_visit(n)
        return self._call('__pde_and' if isinstance(n.op, ast.And) else '__pde_or',
                          [self._thunk(v) for v in n.values])

    def visit_UnaryOp(self, n):
        self.generic_visit(n)
        if isinstance(n.op, ast.Not):
            return self._call('__pde_not', [n.operand])
        return n


class _Counter:
    def __init__(self):
        self.n = 0

    def next(self):
        self.n += 1
        return self.n


def _name(id_, store=False):
    return ast.Name(id=id_, ctx=ast.Store() if store else ast.Load())


def _rewrite_block(stmts, bound_before, ctr):
    """Rewrites the `if` statements of a function-body block (see the comment above).
    bound_before: names bound earlier in the enclosing function (parameters included)."""
    out = []
    bound = list(bound_before)
    for i, st in enumerate(stmts):
        if not isinstance(st, ast.If):
            out.append(st)
            for nme in _stored_names([st]):
                if nme not in bound:
                    bound.append(nme)
            continue
        k = ctr.next()
        cname = '__pde_c%d' % k
        out.append(ast.Assign(targets=[_name(cname, True)], value=st.test))
        rest = stmts[i + 1:]
        if _has_return(st.body) or _has_return(st.orelse):
            # early return: move the rest of the block into both arms
            arms = []
            for tag, blk in (('t', st.body), ('e', st.orelse)):
                full = list(blk) + ([] if _ends_in_return(blk) else _copy.deepcopy(rest))
                if not _ends_in_return(full):
                    full = full + [ast.Return(value=ast.Constant(value=None))]
                # variables the arm rebinds or mutates and that exist already enter as
                # (copied) default arguments: the arm must not see the other arm's writes
                params = [nme for nme in _stored_names(full) if nme in bound]
                fn = ast.FunctionDef(
                    name='__pde_%s%d' % (tag, k),
                    args=ast.arguments(
                        posonlyargs=[], args=[ast.arg(arg=p_) for p_ in params], kwonlyargs=[],
                        kw_defaults=[],
                        defaults=[_ExprRewriter._call('__pde_cp', [_name(p_)]) for p_ in params]),
                    body=_rewrite_block(full, bound, ctr), decorator_list=[], returns=None,
                    type_params=[])
                out.append(fn)
                arms.append(_name(fn.name))
            out.append(ast.Return(value=_ExprRewriter._call('__pde_ifexp',
                                                            [_name(cname)] + arms)))
            return out
        # no return inside: run both arms on copies, merge what they assign
        names = _stored_names(st.body + st.orelse)
        plain = ast.If(test=_name(cname), body=_rewrite_block(st.body, bound, ctr) or [ast.Pass()],
                       orelse=_rewrite_block(st.orelse, bound, ctr))

        def snap():
            return ast.List(elts=[_ExprRewriter._call('__pde_try',
                                                      [_ExprRewriter._thunk(_name(nme))])
                                  for nme in names], ctx=ast.Load())

        def unpack(value):
            return ast.Assign(targets=[ast.Tuple(elts=[_name(nme, True) for nme in names],
                                                 ctx=ast.Store())], value=value)

        sname, aname = '__pde_s%d' % k, '__pde_a%d' % k
        both = []
        if names:
            both.append(ast.Assign(targets=[_name(sname, True)], value=snap()))
            both += _rewrite_block(_copy.deepcopy(st.body), bound, ctr)
            both.append(ast.Assign(targets=[_name(aname, True)], value=snap()))
            both.append(unpack(_ExprRewriter._call('__pde_restore', [_name(sname)])))
            both += _rewrite_block(_copy.deepcopy(st.orelse), bound, ctr)
            both.append(unpack(_ExprRewriter._call(
                '__pde_merge', [_name(cname), _name(aname), snap(),
                                ast.Tuple(elts=[ast.Constant(value=nme) for nme in names],
                                          ctx=ast.Load())])))
        else:
            both.append(ast.Pass())
        out.append(ast.If(test=_ExprRewriter._call('__pde_plain', [_name(cname)]), body=[plain],
                          orelse=both))
        for nme in names:
            if nme not in bound:
                bound.append(nme)
    return out


def _lower_branches(func):
    """A copy of the Python function `func` with its data-dependent branches rewritten, or
    None where there is nothing to rewrite (or no source to rewrite: the code object is
    then traced as it is)."""
    try:
        src = textwrap.dedent(inspect.getsource(func))
        tree = ast.parse(src)
    except (OSError, TypeError, SyntaxError, IndentationError):
        return None
    fdef = next((n for n in tree.body if isinstance(n, ast.FunctionDef)), None)
    if fdef is None or fdef.name != func.__name__:
        return None
    if not any(isinstance(n, (ast.If, ast.IfExp, ast.BoolOp)) or
               (isinstance(n, ast.UnaryOp) and isinstance(n.op, ast.Not)) for n in ast.walk(fdef)):
        return None
    fdef.decorator_list = []
    a = fdef.args
    params = [x.arg for x in a.posonlyargs + a.args + a.kwonlyargs]
    if a.vararg:
        params.append(a.vararg.arg)
    if a.kwarg:
        params.append(a.kwarg.arg)
    fdef.body = _rewrite_block(fdef.body, params, _Counter())
    _ExprRewriter().visit(fdef)
    # (defaults and annotations are re-attached from the original below)
    a.defaults, a.kw_defaults = [], [None] * len(a.kwonlyargs)
    for x in a.posonlyargs + a.args + a.kwonlyargs:
        x.annotation = None
    fdef.returns = None
    mod = ast.Module(body=[fdef], type_ignores=[])
    ast.fix_missing_locations(mod)
    try:
        code = compile(mod, '<pypde_b200.tracing: %s>' % func.__name__, 'exec')
    except (SyntaxError, ValueError, TypeError):
        return None
    return code, fdef.name


def retarget(func, _seen=None):
    """A copy of `func` whose globals resolve numpy/math names to tracing versions
    and helper functions (plain or numba-jitted) to retargeted copies."""
    _seen = {} if _seen is None else _seen
    func = getattr(func, 'py_func', func)        # numba dispatcher -> Python function
    if not isinstance(func, types.FunctionType):
        return func
    if id(func) in _seen:
        return _seen[id(func)]
    g = dict(func.__globals__)
    lowered = _lower_branches(func)
    if lowered is not None:
        # the rewritten source is compiled against the same (retargeted) globals; closure
        # variables of the original become globals of the copy
        g.update(_PDE_HELPERS)
        if func.__closure__:
            for nme, cell in zip(func.__code__.co_freevars, func.__closure__):
                try:
                    g[nme] = cell.cell_contents
                except ValueError:
                    pass
        ns = {}
        exec(lowered[0], g, ns)
        new = ns[lowered[1]]
        new.__defaults__ = func.__defaults__
    else:
        new = types.FunctionType(func.__code__, g, func.__name__, func.__defaults__,
                                 func.__closure__)
    new.__kwdefaults__ = func.__kwdefaults__
    _seen[id(func)] = new
    for name in func.__code__.co_names:
        if name in _PDE_HELPERS:
            continue
        if name not in func.__globals__:
            b = __builtins__ if isinstance(__builtins__, dict) else vars(__builtins__)
            if name in ('abs', 'sum', 'max', 'min') and name in b:
                g[name] = _NP_IDS[id(b[name])]
            continue
        v = func.__globals__[name]
        if id(v) in _NP_IDS:
            g[name] = _NP_IDS[id(v)]
        elif isinstance(v, types.ModuleType) and v.__name__.split('.')[0] in ('numpy', 'math'):
            g[name] = _Proxy(v)
        elif isinstance(v, types.ModuleType) and v.__name__.split('.')[0] not in (
                'numba', 'scipy', 'sys', 'os'):
            g[name] = _UserModuleProxy(v, _seen)
        elif hasattr(v, 'py_func') or isinstance(v, types.FunctionType):
            g[name] = retarget(v, _seen)
    return new


# ---------------------------------------------------------------------------
# code generation
# ---------------------------------------------------------------------------
def _lit(x):
    x = float(x)
    if x != x or x in (float('inf'), float('-inf')):
        raise TraceError('non-finite constant in a user function')
    return repr(x) if ('e' in repr(x) or '.' in repr(x) or 'inf' in repr(x)) else repr(x) + '.'


def emit_body(tape, outputs, out_name, indent='    '):
    """C statements computing `outputs` (list of Sym / float) into out_name[i]."""
    need = set()
    stack = [o.idx for o in outputs if isinstance(o, Sym)]
    while stack:
        i = stack.pop()
        if i in need:
            continue
        need.add(i)
        for a in tape.nodes[i][1]:
            if isinstance(a, Sym):
                stack.append(a.idx)
    lines = []

    def ref(a):
        return 't%d' % a.idx if isinstance(a, Sym) else _lit(a)

    for i, (op, args) in enumerate(tape.nodes):
        if i not in need:
            continue
        if op == 'in':
            expr = '%s[%d]' % (args[0], args[1])
        elif op in _BIN:
            expr = '%s %s %s' % (ref(args[0]), _BIN[op], ref(args[1]))
        elif op == 'neg':
            expr = '-%s' % ref(args[0])
        elif op in _FUN1:
            expr = '%s(%s)' % (op, ref(args[0]))
        elif op in ('pow', 'fmax', 'fmin'):
            expr = '%s(%s, %s)' % (op, ref(args[0]), ref(args[1]))
        elif op in _CMP:
            lines.append('%sconst bool t%d = %s %s %s;' % (indent, i, ref(args[0]), _CMP[op],
                                                           ref(args[1])))
            continue
        elif op in ('and', 'or'):
            lines.append('%sconst bool t%d = %s %s %s;' % (indent, i, ref(args[0]),
                                                           '&&' if op == 'and' else '||',
                                                           ref(args[1])))
            continue
        elif op == 'not':
            lines.append('%sconst bool t%d = !%s;' % (indent, i, ref(args[0])))
            continue
        elif op == 'select':
            expr = '%s ? %s : %s' % (ref(args[0]), ref(args[1]), ref(args[2]))
        else:
            raise TraceError('unknown traced operation %r' % op)
        lines.append('%sconst double t%d = %s;' % (indent, i, expr))
    for k, o in enumerate(outputs):
        lines.append('%s%s[%d] = %s;' % (indent, out_name, k, ref(o)))
    return '\n'.join(lines)


def _inputs(tape, name, shape):
    a = np.empty(shape, dtype=object)
    flat = 0
    for i in np.ndindex(*shape) if isinstance(shape, tuple) else range(shape):
        a[i] = tape.node('in', name, flat)
        flat += 1
    return a


def trace_function(func, kind, ndim, V):
    """Traces a reference-style F / B / S and returns (cuda_source, second_order).

    kind 'F': func(Q) | func(Q, d) | func(Q, dQ, d) -> array of V
    kind 'B': func(Q) | func(Q, d) -> (V, V);   kind 'S': func(Q) -> array of V
    """
    import inspect
    nargs = len(inspect.signature(func).parameters)
    f = retarget(func)
    second_order = kind == 'F' and nargs == 3
    nout = V * V if kind == 'B' else V
    bodies = []
    dirs = range(ndim) if (kind != 'S' and nargs >= 2) else [0]
    for d in dirs:
        tape = Tape()
        Q = _inputs(tape, 'q', (V, ))
        if kind == 'F':
            dQ = _inputs(tape, 'dq', (ndim, V))
            res = f(Q) if nargs == 1 else (f(Q, d) if nargs == 2 else f(Q, dQ, d))
        elif kind == 'B':
            res = f(Q) if nargs == 1 else f(Q, d)
        else:
            res = f(Q)
        if res is None:
            raise TraceError('the function returned None (device-style function?)')
        res = _obj(res)
        if res.size != nout:
            raise TraceError('%s returned %d values, expected %d' % (kind, res.size, nout))
        bodies.append(emit_body(tape, list(res.ravel()), 'out'))
    sig = {'F': '(double *out, const double *q, const double *dq, int d)',
           'B': '(double *out, const double *q, int d)',
           'S': '(double *out, const double *q)'}[kind]
    src = ['// generated by pypde_b200.tracing from %s' % getattr(func, '__name__', '?'),
           'extern "C" __device__ void user_%s%s {' % (kind, sig)]
    if len(bodies) == 1:
        src.append(bodies[0])
    else:
        src.append('  switch (d) {')
        for d, b in zip(dirs, bodies):
            src.append('  case %d: {' % d)
            src.append(b)
            src.append('  } break;')
        src.append('  }')
    src.append('}')
    return '\n'.join(src) + '\n', second_order
