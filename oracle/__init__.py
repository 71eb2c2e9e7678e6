"""CPU oracle for the Subspace-Multimodal-Learning hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package (``subspace-multimodal-learning_b200``); only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may use it, and there only as the checker / CPU baseline.

Each function is an independent torch-CPU restatement (no einops, no
``F.grid_sample``, no ``nn.Module`` from the reference) of the algorithm at the
cited ``/root/reference`` file:line.  The reference ships no golden vectors or
tests, so the oracle is pinned against outputs of the *imported
reference itself*: ``tests/golden/make_golden.py`` (run in the authoring
container, where ``/root/reference`` exists) writes small seeded fixtures to
``tests/golden/*.pt`` and ``tests/test_oracle_golden.py`` checks every oracle
function against them.  Third-party arithmetic: the live ``NystromAttention``
import is the un-vendored, unpinned PyPI package ``nystrom-attention``
(lucidrains); the reference vendors two verbatim copies
(models/NystromAttention.py:20-157, models/cmta_utils.py:147-281) and the
fixtures are generated from the first.
"""
from . import nystrom, deform, heads, loss, modulation, pwl, cmta  # noqa: F401
