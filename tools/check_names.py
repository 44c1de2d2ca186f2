"""Poor man's pyflakes (not installed here): report names that are read somewhere in a module but never bound anywhere in it
(assignment, def, class, import, argument, comprehension / loop / with / except target) and are not builtins.
    python tools/check_names.py yolov5m_b200/*.py bench.py"""
import ast
import builtins
import sys


def check(path):
    tree = ast.parse(open(path).read(), path)
    bound, used = set(dir(builtins)) | {"__file__", "__name__", "__doc__"}, []
    for node in ast.walk(tree):
        if isinstance(node, ast.Name):
            (bound.add if isinstance(node.ctx, (ast.Store, ast.Del)) else lambda n: used.append((n, node.lineno)))(node.id)
        elif isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            bound.add(node.name)
            if not isinstance(node, ast.ClassDef):
                a = node.args
                for x in a.args + a.kwonlyargs + a.posonlyargs + ([a.vararg] if a.vararg else []) + ([a.kwarg] if a.kwarg else []):
                    bound.add(x.arg)
        elif isinstance(node, ast.Lambda):
            a = node.args
            for x in a.args + a.kwonlyargs + a.posonlyargs + ([a.vararg] if a.vararg else []) + ([a.kwarg] if a.kwarg else []):
                bound.add(x.arg)
        elif isinstance(node, (ast.Import, ast.ImportFrom)):
            for al in node.names:
                bound.add((al.asname or al.name).split(".")[0])
        elif isinstance(node, ast.ExceptHandler) and node.name:
            bound.add(node.name)
    bad = sorted({(n, l) for n, l in used if n not in bound})
    for n, l in bad:
        print(f"{path}:{l}: name '{n}' is never bound in this module")
    return len(bad)


if __name__ == "__main__":
    sys.exit(1 if sum(check(p) for p in sys.argv[1:]) else 0)
