"""Extracts the reference's own known-answer positions into a JSON fixture.

The reference (``/root/reference``) does not exist on the GPU box, and its engine cannot be built
anywhere in this environment, so the only golden vectors for the hot path are the
pictured positions in ``/root/reference/tests/test_connect.py`` and ``tests/test_bounce.py``.
This script walks those files' ASTs, feeds every board picture through the reference's OWN
``parse()`` helper, and records -- per test function, in order -- the grid, the side to move, the
selected move, the exhaustive target set (Bounce) and the literal rewards / JSON dicts the test
asserts.  Nothing is computed by this repo's engine or oracle: the fixture is purely the
reference's expectations.

Run (in the build container, where /root/reference exists):
    python tests/golden/make_golden.py
"""
from __future__ import annotations

import ast
import importlib.util
import json
import os
import sys

REF = "/root/reference/tests"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))


def _load(name):
    # the test modules import `simulator.game.*` at module level; any importable stand-in will do,
    # only their pure-Python parse() is used here
    sys.path.insert(0, os.path.join(ROOT, "oracle", "pyapi"))
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _literal(node):
    return ast.literal_eval(node)


def _walk_test(fn: ast.FunctionDef):
    """Yields ('picture', str) / ('reward', list) / ('config', [h,w,k]) / ('json', kind, dict)."""
    for node in ast.walk(fn):
        pass
    events = []

    class V(ast.NodeVisitor):
        def visit_Call(self, node):
            f = node.func
            name = f.id if isinstance(f, ast.Name) else (f.attr if isinstance(f, ast.Attribute) else None)
            if name == "assert_state" and node.args and isinstance(node.args[0], ast.Constant):
                events.append((node.lineno, "picture", node.args[0].value))
            elif name == "Config" and node.args and all(isinstance(a, ast.Constant) for a in node.args):
                events.append((node.lineno, "config", [a.value for a in node.args]))
            elif name == "assert_array_equal" and len(node.args) == 2 and isinstance(node.args[1], ast.List):
                src = ast.unparse(node.args[0])
                if src.endswith(".reward"):
                    events.append((node.lineno, "reward", _literal(node.args[1])))
            self.generic_visit(node)

        def visit_Compare(self, node):
            left = ast.unparse(node.left)
            if len(node.comparators) == 1:
                right = node.comparators[0]
                if left.endswith(".reward.tolist()") and isinstance(right, ast.List):
                    events.append((node.lineno, "reward", _literal(right)))
                elif left.endswith(".to_json()") and isinstance(right, ast.Dict):
                    events.append((node.lineno, "json:" + left.split(".")[0], _literal(right)))
                elif left == "len(state.actions)" and isinstance(right, ast.Constant):
                    events.append((node.lineno, "n_actions", right.value))
            self.generic_visit(node)

        def visit_Assert(self, node):
            if ast.unparse(node.test) == "state.has_ended":
                events.append((node.lineno, "has_ended", True))
            self.generic_visit(node)

    V().visit(fn)
    events.sort(key=lambda e: e[0])
    return [(k, v) for _, k, v in events]


def extract(module_name, game):
    mod = _load(module_name)
    tree = ast.parse(open(os.path.join(REF, module_name + ".py")).read())
    out = {}
    for fn in tree.body:
        if not (isinstance(fn, ast.FunctionDef) and fn.name.startswith("test_")):
            continue
        rec = {"steps": [], "final": {}, "json": {}}
        for kind, val in _walk_test(fn):
            if kind == "picture":
                if game == "connect":
                    grid, column, player = mod.parse(val)
                    rec["steps"].append({"grid": grid.tolist(), "column": column, "player": int(player)})
                else:
                    grid, selections, targets = mod.parse(val)
                    # same inference as the reference's assert_state (test_bounce.py:43-48,63-76)
                    player, source, target = 0, None, None
                    if selections:
                        [(sx, sy)] = set(selections) - set(targets)
                        source = [sx, sy]
                        if grid[:sy].sum() > 0:
                            player = 1
                        picked = [s for s in selections if s in targets]
                        target = list(picked[0]) if picked else None
                    rec["steps"].append({
                        "grid": grid.tolist(), "player": player, "source": source,
                        "targets": sorted([list(t) for t in targets]) if selections else None,
                        "target": target,
                    })
            elif kind == "config":
                rec["config"] = val
            elif kind in ("reward", "n_actions", "has_ended"):
                rec["final"][kind] = val
            elif kind.startswith("json:"):
                rec["json"][kind[5:]] = val
        out[fn.name] = rec
    return out


def main():
    fixture = {
        "_source": "extracted from /root/reference/tests/test_connect.py and test_bounce.py by tests/golden/make_golden.py",
        "connect": extract("test_connect", "connect"),
        "bounce": extract("test_bounce", "bounce"),
    }
    path = os.path.join(HERE, "reference_positions.json")
    with open(path, "w") as f:
        json.dump(fixture, f, indent=1, sort_keys=True)
    n = sum(len(t["steps"]) for g in ("connect", "bounce") for t in fixture[g].values())
    print(f"wrote {path}: {n} pictured positions")


if __name__ == "__main__":
    main()
