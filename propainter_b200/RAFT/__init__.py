from .raft import RAFT  # noqa: F401  (same import surface as the reference's `from RAFT import RAFT`)
