"""CPU oracle for the MSMFormer hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional (weights-in-a-dict) restatement, in plain PyTorch-CPU / numpy, of the reference
algorithm on the path SURVEY.md §8(a) lists. Each function cites the reference file:line it
follows (paths relative to the reference checkout, YoungSean/UnseenObjectsWithMeanShift @ d1c8487).

Who may import this package: ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` - as the checker or the timed CPU baseline, never on
the product path. ``unseenobjectswithmeanshift_b200`` must not import it (tests/test_boundary.py
greps for that) and fails loudly when its CUDA library is missing.

Parity pinning: the reference ships no stored vectors for this path (its only known-answer
recipe is pixel_decoder/ops/test.py). The oracle is therefore pinned against outputs of the
reference's own modules, imported by file path in the authoring container and committed as
``tests/golden/*.npz`` together with the generator (``tests/golden/make_golden.py``);
``tests/test_oracle_golden.py`` checks every function here against them.

The reference is pure PyTorch on this path except the MSDeformAttn CUDA op, which cannot run
on a CPU (ops/src/ms_deform_attn.h:43) and does not compile against torch 2.11 as shipped
(ms_deform_attn_cuda.cu:69,139) - there is no ``oracle/_ref`` binary; see DESIGN.md.
"""
