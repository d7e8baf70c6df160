"""Import hook that runs the UNMODIFIED reference sources under the taichi stand-in.

TEST INFRASTRUCTURE ONLY (oracle tier 1).  `install()` maps the package name
`thinshelllab` onto /root/reference/code (as the reference's pyproject.toml:29-30
does), puts the stand-in `taichi`, `cupy`, `cupyx`, `matplotlib` modules on
sys.path and compiles every reference module through a small AST pass that
restores two Taichi semantics plain Python lacks:

  1. `ti.atomic_add/max/min(lvalue, v)` take an l-value and return the old value
     (e.g. code/engine/BaseScene.py:797, code/engine/geometry.py:112,152,
     code/engine/sparse_solver.py:35,38);
  2. `a = field[i]` copies the value in kernel scope (numpy would alias).

No reference source text is copied into this repository; the files are read from
/root/reference at run time, which is why this only works in the build container.
"""
import ast
import importlib.abc
import importlib.util
import os
import sys

REF_CODE = os.environ.get("TSL_REFERENCE_CODE", "/root/reference/code")
_HERE = os.path.dirname(os.path.abspath(__file__))
_ATOMICS = {"atomic_add": "add", "atomic_max": "max", "atomic_min": "min"}


def _is_atomic(node):
    return (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute)
            and node.func.attr in _ATOMICS and isinstance(node.func.value, ast.Name)
            and node.func.value.id == "ti")


def _emu(name):
    return ast.Attribute(value=ast.Name(id="_ti_emu_", ctx=ast.Load()), attr=name, ctx=ast.Load())


class _Rewrite(ast.NodeTransformer):
    def _name_atomic(self, call, assign_to=None):
        """atomic on a plain local: returns replacement statements"""
        op = _ATOMICS[call.func.attr]
        tgt = call.args[0]
        val = self.visit(call.args[1])
        out = []
        if assign_to is not None:
            out.append(ast.Assign(targets=[assign_to], value=ast.Name(id=tgt.id, ctx=ast.Load())))
        out.append(ast.Assign(
            targets=[ast.Name(id=tgt.id, ctx=ast.Store())],
            value=ast.Call(func=_emu("_emu_combine"),
                           args=[ast.Constant(op), ast.Name(id=tgt.id, ctx=ast.Load()), val], keywords=[])))
        return out

    def visit_Expr(self, node):
        if _is_atomic(node.value) and isinstance(node.value.args[0], ast.Name):
            return self._name_atomic(node.value)
        return self.generic_visit(node)

    def visit_Assign(self, node):
        if (_is_atomic(node.value) and isinstance(node.value.args[0], ast.Name)
                and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name)):
            return self._name_atomic(node.value, node.targets[0])
        node = self.generic_visit(node)
        node.value = self._val(node.value)
        return node

    def _val(self, v):
        if isinstance(v, ast.Tuple):
            v.elts = [self._val(e) for e in v.elts]
            return v
        if isinstance(v, (ast.Subscript, ast.Name, ast.Attribute)):
            return ast.Call(func=_emu("_emu_val"), args=[v], keywords=[])
        return v

    def visit_Call(self, node):
        node = self.generic_visit(node)
        if _is_atomic(node):
            tgt = node.args[0]
            if isinstance(tgt, ast.Subscript):
                return ast.Call(func=_emu("_emu_atomic"),
                                args=[ast.Constant(_ATOMICS[node.func.attr]), tgt.value, tgt.slice, node.args[1]],
                                keywords=[])
            raise SyntaxError("atomic on unsupported l-value: " + ast.dump(tgt))
        return node


class _Loader(importlib.abc.Loader):
    def __init__(self, path, is_pkg):
        self.path, self.is_pkg = path, is_pkg

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        src_path = os.path.join(self.path, "__init__.py") if self.is_pkg else self.path
        with open(src_path, "r") as f:
            src = f.read()
        tree = ast.parse(src, filename=src_path)
        tree = _Rewrite().visit(tree)
        imp = ast.Import(names=[ast.alias(name="taichi", asname="_ti_emu_")])
        # keep `from __future__` first if present
        pos = 0
        while pos < len(tree.body) and isinstance(tree.body[pos], ast.ImportFrom) and tree.body[pos].module == "__future__":
            pos += 1
        if tree.body and isinstance(tree.body[0], ast.Expr) and isinstance(getattr(tree.body[0], "value", None), ast.Constant):
            pos = max(pos, 1)
        tree.body.insert(pos, imp)
        ast.fix_missing_locations(tree)
        code = compile(tree, src_path, "exec")
        module.__file__ = src_path
        exec(code, module.__dict__)


class _Finder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname != "thinshelllab" and not fullname.startswith("thinshelllab."):
            return None
        rel = fullname.split(".")[1:]
        base = os.path.join(REF_CODE, *rel)
        if os.path.isdir(base):
            spec = importlib.util.spec_from_loader(fullname, _Loader(base, True), is_package=True)
            spec.submodule_search_locations = [base]
            return spec
        if os.path.isfile(base + ".py"):
            return importlib.util.spec_from_loader(fullname, _Loader(base + ".py", False))
        return None


_installed = False


def install():
    """Idempotent.  After this, `import thinshelllab.task_scene.Scene_bouncing` runs the
    reference source from REF_CODE under emulation."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(REF_CODE):
        raise RuntimeError(f"reference sources not found at {REF_CODE} (tier-1 oracle only runs in the build container)")
    sys.dont_write_bytecode = True
    sys.path.insert(0, _HERE)                          # stand-in `taichi`
    sys.path.insert(1, os.path.join(_HERE, "stubs"))   # cupy / cupyx / matplotlib stand-ins
    sys.meta_path.insert(0, _Finder())
    import numpy as np
    if not hasattr(np, "product"):
        np.product = np.prod   # code/engine/model_elastic_offset.py:27 needs NumPy < 2
    _installed = True
