"""CPU oracle for the denoising-sampler hot path (TEST INFRASTRUCTURE ONLY).

A plain PyTorch-CPU / numpy fp32 restatement of the reference algorithm
(entheeb/A-Multimodal-Diffusion-Based-Model-for-Point-Cloud-Completion), each
function citing the reference file:line it follows.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it, and there only as the checker or the timed CPU baseline --
never as (part of) the product path.

Pinning: the reference ships NO tests or golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the
reference itself executed in the authoring container
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``; checked by
``tests/test_oracle_golden.py``).
"""
