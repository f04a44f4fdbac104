"""CPU oracle for the GConv-Adapter hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker or the timed CPU baseline.
The product path (``gconv_adapter_b200``) never imports this package and fails
loudly when its CUDA library is missing.

PARITY UNPINNED: the arithmetic of the reference's hot path
lives in torch-geometric (``torch-geometric==2.5.3``,
/root/reference/requirements/mpnn_inductive_environment.yml:102, and ``==1.7.2``,
/root/reference/requirements/gtn_transductive_environment.yaml:91), which is neither
vendored under /root/reference nor installed in this image, and the reference ships no
tests, golden vectors or fixtures.  ``pyg_restated.py`` therefore restates the published
``GCNConv`` / ``gcn_norm`` / ``add_remaining_self_loops`` algorithm op for op and is anchored
on (a) the reference's own call sites (src/finetune/gconv_adapter.py:40-41,73-78,92),
(b) the reference's *real* glue class executed in this container over the restated conv
(``pyg_shim.py`` + ``tests/golden/make_golden.py``), (c) the two in-repo normalisation
formulas (src/layers/inductive/gcn_conv.py:36-56,
src/dataset/transductive/data_utils.py:175-183) and (d) hand-computed tiny graphs.
"""
