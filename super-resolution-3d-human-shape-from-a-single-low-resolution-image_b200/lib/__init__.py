"""Mirror of the reference's ``lib`` package for the reconstruction hot path (drop-in names)."""
