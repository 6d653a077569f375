#!/usr/bin/env python
"""Render INTEGRATION.md from integration/INTEGRATION.md.in with the patch and the binding embedded verbatim
(tests/test_abi.py checks that the document is in step with both files)."""
import os

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def render() -> str:
    t = open(os.path.join(HERE, "INTEGRATION.md.in")).read()
    t = t.replace("@@PATCH@@", open(os.path.join(HERE, "driver_gmx.patch")).read())
    return t.replace("@@BRIDGE@@", open(os.path.join(HERE, "gnumap_gmx_bridge.cpp")).read())


if __name__ == "__main__":
    open(os.path.join(ROOT, "INTEGRATION.md"), "w").write(render())
