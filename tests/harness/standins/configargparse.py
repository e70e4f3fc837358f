"""Harness stand-in for the ``configargparse`` package (not installed offline).

Only what run_nerf.py:33-186 uses: ``ArgumentParser`` whose ``add_argument`` accepts ``is_config_file=True`` and
whose ``parse_args`` merges ``key = value`` lines of the config file under the command line.  TEST HARNESS ONLY --
it sits on the harness' sys.path so that the reference's run_nerf.py can be executed UNCHANGED."""
import argparse
import sys


class ArgumentParser(argparse.ArgumentParser):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._config_dests = []

    def add_argument(self, *names, **kw):
        if kw.pop("is_config_file", False):
            act = super().add_argument(*names, **kw)
            self._config_dests.append(act.dest)
            return act
        return super().add_argument(*names, **kw)

    def parse_args(self, args=None, namespace=None):
        argv = list(sys.argv[1:] if args is None else args)
        pre, _ = super().parse_known_args(argv)
        extra = []
        for dest in self._config_dests:
            path = getattr(pre, dest, None)
            if not path:
                continue
            with open(path) as fh:
                for line in fh:
                    line = line.split("#", 1)[0].strip()
                    if not line or "=" not in line:
                        continue
                    key, val = (s.strip() for s in line.split("=", 1))
                    flag = "--" + key
                    if val.lower() == "true":
                        extra.append(flag)
                    elif val.lower() not in ("false", "none", ""):
                        extra += [flag, val]
        return super().parse_args(extra + argv, namespace)   # the command line wins over the file
