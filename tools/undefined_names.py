"""Undefined-name check (no linter is installed in this image): every Name that is loaded in a function or at module
level must be bound somewhere in an enclosing scope, be a module global, or a builtin.

    python tools/undefined_names.py joligen_b200/*.py bench.py

Conservative: star imports / exec / globals() tricks make it give up on the file.  tests/test_host_logic.py runs it
over the package so that a forgotten import in a rarely taken branch (the kind of error no GPU test reaches) fails
on the CPU.
"""
import ast
import builtins
import sys


class _Scope:
    def __init__(self, parent=None, is_class=False):
        self.parent = parent
        self.is_class = is_class
        self.bound = set()
        self.loads = []  # (name, lineno)


def _targets(node, scope):
    for n in ast.walk(node):
        if isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            scope.bound.add(n.id)


class _Checker(ast.NodeVisitor):
    def __init__(self):
        self.module = _Scope()
        self.scope = self.module
        self.scopes = [self.module]

    def _push(self, is_class=False):
        s = _Scope(self.scope, is_class)
        self.scopes.append(s)
        self.scope = s
        return s

    def _pop(self):
        self.scope = self.scope.parent

    def visit_Import(self, node):
        for a in node.names:
            self.scope.bound.add((a.asname or a.name).split(".")[0])

    def visit_ImportFrom(self, node):
        for a in node.names:
            if a.name == "*":
                raise NotImplementedError("star import")
            self.scope.bound.add(a.asname or a.name)

    def _function(self, node):
        if not isinstance(node, ast.Lambda):
            self.scope.bound.add(node.name)
            for d in node.decorator_list:
                self.visit(d)
            if node.returns:
                self.visit(node.returns)
        args = node.args
        for d in args.defaults + [d for d in args.kw_defaults if d is not None]:
            self.visit(d)
        s = self._push()
        for a in args.posonlyargs + args.args + args.kwonlyargs + [x for x in (args.vararg, args.kwarg) if x]:
            s.bound.add(a.arg)
        body = node.body if isinstance(node.body, list) else [node.body]
        for b in body:
            self.visit(b)
        self._pop()

    visit_FunctionDef = visit_AsyncFunctionDef = visit_Lambda = _function

    def visit_ClassDef(self, node):
        self.scope.bound.add(node.name)
        for d in node.decorator_list + node.bases + [k.value for k in node.keywords]:
            self.visit(d)
        self._push(is_class=True)
        for b in node.body:
            self.visit(b)
        self._pop()

    def _comprehension(self, node):
        s = self._push()
        for g in node.generators:
            _targets(g.target, s)
        for g in node.generators:
            self.visit(g.iter)
            for c in g.ifs:
                self.visit(c)
        for f in ("elt", "key", "value"):
            if hasattr(node, f):
                self.visit(getattr(node, f))
        self._pop()

    visit_ListComp = visit_SetComp = visit_DictComp = visit_GeneratorExp = _comprehension

    def visit_Global(self, node):
        for n in node.names:
            self.module.bound.add(n)
            self.scope.bound.add(n)

    visit_Nonlocal = visit_Global

    def visit_ExceptHandler(self, node):
        if node.name:
            self.scope.bound.add(node.name)
        self.generic_visit(node)

    def visit_Name(self, node):
        if isinstance(node.ctx, ast.Load):
            self.scope.loads.append((node.id, node.lineno))
        else:
            self.scope.bound.add(node.id)

    def visit_NamedExpr(self, node):
        # (binds in the enclosing function scope even inside a comprehension; good enough: bind here and upwards)
        s = self.scope
        while s is not None:
            s.bound.add(node.target.id)
            s = s.parent
        self.visit(node.value)

    def visit_MatchAs(self, node):
        if node.name:
            self.scope.bound.add(node.name)
        self.generic_visit(node)


def check_source(src, filename="<src>"):
    """-> list of (lineno, name) loaded but bound nowhere visible."""
    tree = ast.parse(src, filename)
    c = _Checker()
    c.visit(tree)
    known = set(dir(builtins)) | {"__file__", "__name__", "__doc__", "__path__", "__spec__", "__class__"}
    out = []
    for s in c.scopes:
        for name, line in s.loads:
            t, found = s, False
            first = True
            while t is not None:
                # class bodies are not visible from nested functions
                if (first or not t.is_class) and name in t.bound:
                    found = True
                    break
                first = False
                t = t.parent
            if not found and name not in known:
                out.append((line, name))
    return sorted(set(out))


def main(paths):
    bad = 0
    for p in paths:
        try:
            res = check_source(open(p).read(), p)
        except NotImplementedError as e:
            print("%s: skipped (%s)" % (p, e))
            continue
        for line, name in res:
            print("%s:%d: undefined name '%s'" % (p, line, name))
            bad += 1
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
