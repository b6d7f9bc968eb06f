"""CPU oracle for the CP linear-transformer hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the timed CPU
baseline.  The product package never imports this module and fails loudly when
its CUDA library is missing.

PINNED WHERE THE REFERENCE'S OWN CODE RUNS: ``tests/golden/make_ref_golden.py``
imports the reference's model / training files UNMODIFIED from ``/root/reference``
(in the build container) and executes them — embeddings, positional encoding,
in_linear, the teacher-forced and recurrent call protocols, the 6 heads, masked
CE, the numpy samplers, critic / actor value paths, ``PPO`` and ``DQN`` class
arithmetic, the two reward read-outs — with only ``fast_transformers`` supplied
by ``ft_oracle.py``; ``tests/test_ref_golden.py`` holds this oracle to those
committed vectors (``tests/golden/ref_model.npz``, ``ref_rl.npz``).

PARITY UNPINNED FOR THE ENCODER INTERNALS: the arithmetic inside the encoder
lives in the third-party package
``pytorch-fast-transformers==0.4.0`` (reference ``requirements.txt:54``), which
is neither vendored under ``/root/reference`` nor installable offline, and the
reference ships no tests, golden vectors or checkpoints.
The oracle therefore restates the *published* fast_transformers 0.4.0 algorithm
 and the reference's own call sites, and is cross-checked
three independent ways (quadratic masked form == cumulative-sum form ==
recurrent form == the C clone of ``causal_product_cpu``), but it cannot be
checked against outputs of the real dependency in this container.
"""

from . import ft_oracle, model_oracle, sampling_oracle, rl_oracle  # noqa: F401
