"""CPU oracle for the DDPO hot path (TEST INFRASTRUCTURE ONLY).

This package restates, in NumPy / torch-CPU, the arithmetic of the reference's
denoise-sample -> reward -> PPO-update path (jannerm/ddpo @ f0b6ca7) and of the
pinned third-party packages it calls (jax 0.4.8 threefry PRNG, diffusers 0.12.1
Flax U-Net, optax 0.1.5 clip+adamw).  Only ``tests/``, ``__graft_entry__.smoke``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it;
the product package ``ddpo_b200`` never does.

Pinning status: the reference ships no tests, fixtures or golden vectors for
this path and JAX/Flax/diffusers are not installable here (no network), so the
U-Net / scheduler / PPO restatement is **parity unpinned** against a live JAX
run.  The PRNG restatement *is* pinned: Random123 threefry2x32 known-answer
vectors and the published JAX values for ``PRNGKey(0)``, ``split`` and
``normal`` (see ``oracle/threefry.py`` and ``tests/test_oracle_prng.py``).
"""
