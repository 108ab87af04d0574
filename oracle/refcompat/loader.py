"""Ref-compat harness: execute the reference's OWN files in this container.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (nxblink_b200/)
imports this module.  It can only work where ``/root/reference`` exists (the
build container); on the GPU box it raises ``ReferenceUnavailable`` and the
tests that need it are skipped -- the vectors they produced are committed
under ``tests/golden/`` together with the script that made them
(``oracle/refcompat/make_golden.py``).

What it does:
  * bypasses ``nxblink/__init__.py`` (needs the removed ``numpy.testing.Tester``)
    by registering an empty package object whose ``__path__`` points at the
    reference directory, so ``import nxblink.raoteh`` executes the reference's
    unmodified source files where they lie;
  * puts stand-ins for exactly the missing third-party symbols on ``sys.path``
    (``oracle/refcompat/shims``: nxmctree.sampling.sample_history,
    nxctmctree.trajectory.{Event,LightTrajectory,get_node_to_tm},
    nxctmctree.navigation.partition_nodes, algopy->numpy, StringIO->io).

No reference source is copied into this repository.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('NXBLINK_REFERENCE_ROOT', '/root/reference')
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shims')


class ReferenceUnavailable(RuntimeError):
    pass


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'nxblink'))


def load():
    """Return the (stub) ``nxblink`` package whose submodules are the reference's files."""
    if not available():
        raise ReferenceUnavailable(
                'reference tree not found at %s' % REFERENCE_ROOT)
    if _SHIMS not in sys.path:
        sys.path.insert(0, _SHIMS)
    pkg = sys.modules.get('nxblink')
    if pkg is None or not getattr(pkg, '_refcompat', False):
        pkg = types.ModuleType('nxblink')
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, 'nxblink')]
        pkg._refcompat = True
        sys.modules['nxblink'] = pkg
    for name in ('model', 'util', 'trajectory', 'navigation', 'poisson',
            'chunking', 'raoteh', 'summary', 'em', 'maxlikelihood',
            'toymodel', 'toydata', 'compound', 'denseutil'):
        setattr(pkg, name, importlib.import_module('nxblink.' + name))
    return pkg
