"""Drop-in for the reference's ``pytorch/bp`` package (same module and symbol names),
backed by libldpc_b200.so.  Import with ``ldpc-sims_b200/`` on sys.path, exactly as the
reference's scripts are run from ``pytorch/``."""
