"""The mirror modules offer the reference's names with the reference's signatures (live: reads the reference's sources
where /root/reference is mounted; nothing is imported from it).  Every module-level function, class, method and
module-level constant of the eight modules on and around the path must exist under the same name; shared functions
take the same positional arguments in the same order with the same defaults (mutable defaults may be ``None`` here);
plotting / animation / download entry points exist and say that they are out of scope."""
import ast
import importlib
import inspect
import os

import pytest

from oracle import refshim

pytestmark = pytest.mark.skipif(not refshim.reference_available(), reason="reference not mounted")

MODULES = {"weatherTLKT": "model", "simulatorTLKT": "model", "utils": "solver", "worker": "solver",
           "master_node": "solver", "forest": "solver", "master": "solver", "isochrones": "isochrones"}
# module-level names of the reference that are imports, script leftovers or plotting configuration, not API
IGNORED_GLOBALS = {"mpl", "rc", "plt", "animation", "Basemap", "rgi", "sys", "font"}


def _reference_api(path):
    tree = ast.parse(open(path).read())
    funcs, consts = {}, set()
    for node in tree.body:
        if isinstance(node, ast.ClassDef):
            funcs[node.name] = None
            for m in node.body:
                if isinstance(m, ast.FunctionDef):
                    funcs[node.name + "." + m.name] = m.args
        elif isinstance(node, ast.FunctionDef):
            funcs[node.name] = node.args
        elif isinstance(node, ast.Assign):
            consts.update(t.id for t in node.targets if isinstance(t, ast.Name))
    return funcs, consts - IGNORED_GLOBALS


@pytest.mark.parametrize("module", sorted(MODULES))
def test_mirror_offers_the_reference_names_and_signatures(module):
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # (the reference's sources hold invalid escape sequences in docstrings)
        funcs, consts = _reference_api(os.path.join(refshim.REFERENCE_ROOT, "code", MODULES[module], module + ".py"))
    mirror = importlib.import_module("iboat_pmcts_b200." + module)
    missing = [c for c in consts if not hasattr(mirror, c)]
    assert not missing, "constants missing from %s: %s" % (module, missing)
    for name, args in funcs.items():
        obj, stub = mirror, False
        for part in name.split("."):
            assert hasattr(obj, part), "%s.%s is missing" % (module, name)
            obj = getattr(obj, part)
            stub = "." in getattr(obj, "__name__", "")  # an out-of-scope entry point is named "Class.method" / "module.function"
            if stub:
                break
        if args is None or stub:
            continue  # a class, or an out-of-scope entry point (checked below)
        want = [a.arg for a in args.args if a.arg not in ("self", "cls")]
        sig = inspect.signature(obj)
        got = [p for p in sig.parameters if p not in ("self", "cls")]
        assert got[:len(want)] == want, "%s.%s takes %s, the reference %s" % (module, name, got, want)
        for pname, ref_default in zip(want[len(want) - len(args.defaults):], args.defaults):
            try:
                ref_value = ast.literal_eval(ref_default)
            except (ValueError, SyntaxError):
                continue  # an expression (e.g. -3.5 + 360, range(0, 21), dict()): compared by the drop-in tests
            mine = sig.parameters[pname].default
            assert mine == ref_value or (mine is None and ref_value in ([], {})), (module, name, pname, mine, ref_value)


def test_out_of_scope_entry_points_say_so():
    from iboat_pmcts_b200 import forest, isochrones, master, simulatorTLKT, utils, weatherTLKT
    for f in (weatherTLKT.Weather.plotQuiver, weatherTLKT.Weather.download, simulatorTLKT.Simulator.play_scenario,
              master.MasterTree.plot_tree_uct, forest.play_multiple_scenarios, forest.download_scenarios,
              isochrones.plot_comparision, utils.Player):
        with pytest.raises(NotImplementedError, match="outside the scope"):
            f()
