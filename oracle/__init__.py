"""oracle/ -- CPU restatement of the reference's per-pixel segmentation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``rnd_semantic_segmentation_b200/`` may
import this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker / the timed CPU baseline, never as the product path.

Where the arithmetic lives: the reference (taintpro98/rnd-semantic-segmentation)
is 100 % Python and bottoms out in PyTorch (pinned ``pytorch=1.7.1`` /
``torchvision=0.8.2`` in the reference's ``environment.yml:42,48``; this image
has torch 2.11).  ``torch_oracle`` restates the reference modules/functions with
the same torch CPU calls the reference makes (each function cites the
reference file:line it follows); ``np_oracle`` restates the underlying ATen
algorithms (align-corners bilinear, log-softmax/NLL, softmax->first-max,
bincount confusion matrix) in numpy so the two can be checked against each
other.

Parity pinning: the reference ships NO tests, golden vectors or fixtures for
this path, so the oracle is pinned against outputs of the
reference's own files imported by path in the authoring container
(``oracle/ref_loader.py`` + ``oracle/gen_golden.py`` -> ``tests/golden/*.npz``),
plus the survey's anchor known-answer value (seed-0 config-1 loss = 4.04814).
"""
