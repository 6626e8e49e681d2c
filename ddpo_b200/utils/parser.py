"""Config loader -- mirrors the behaviour of the reference's ``ddpo/utils/parser.py:71-164``:
``Parser().parse_args(experiment)`` merges ``base[experiment]`` < ``<dataset>["common"]`` < ``<dataset>[experiment]``
< command-line ``--key value`` overrides (type-coerced from the existing value), resolves lazy ``f:`` strings
against the final namespace and adds the process index to the seed (:174-179).  stdlib only (no ``tap``)."""
import importlib
import os
import sys
from types import SimpleNamespace


def _coerce(old, val):
    if isinstance(val, str):
        if val == "None":
            return None
        if isinstance(old, bool):
            return val.lower() in ("1", "true", "yes")
        if isinstance(old, int) and not isinstance(old, bool):
            return int(val)
        if isinstance(old, float):
            return float(val)
        if old is None:
            for cast in (int, float):
                try:
                    return cast(val)
                except ValueError:
                    pass
    return val


class Parser:
    config = "ddpo_b200.config.base"
    dataset = "compressed_animals"

    def parse_args(self, experiment, argv=None):
        argv = list(sys.argv[1:] if argv is None else argv)
        overrides = {}
        i = 0
        while i < len(argv):
            assert argv[i].startswith("--"), f"expected --key value, got {argv[i]}"
            overrides[argv[i][2:]] = argv[i + 1]
            i += 2
        config = overrides.pop("config", self.config)
        dataset = overrides.pop("dataset", self.dataset).replace("-", "_")
        mod = importlib.import_module(config)
        params = dict(mod.base[experiment])
        ds = getattr(mod, dataset)
        params.update(ds.get("common", {}))
        params.update(ds.get(experiment, {}))
        for k, v in overrides.items():
            assert k in params, f"[ utils/parser ] unknown key {k}"
            params[k] = _coerce(params[k], v)
        params.update(config=config, dataset=dataset)
        ns = SimpleNamespace(**params)
        for k, v in list(vars(ns).items()):
            if isinstance(v, str) and v.startswith("f:"):
                setattr(ns, k, eval("f'" + v[2:] + "'", {}, vars(ns)))  # lazy f-string against the namespace
        if getattr(ns, "seed", None) is not None:
            ns.seed = ns.seed + int(os.environ.get("RANK", "0"))
        logbase = getattr(ns, "logbase", None)
        if logbase:
            for k in ("loadpath", "savepath", "modelpath"):
                if isinstance(getattr(ns, k, None), str) and getattr(ns, k):
                    setattr(ns, k, os.path.join(logbase, getattr(ns, k)))
        ns._dict = {k: v for k, v in vars(ns).items() if not k.startswith("_")}
        return ns
