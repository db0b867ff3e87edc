"""Batch containers and dataset plugins: `base` (Data, Edges, DataLoader, collater), `synthetic` (seeded BASELINE
shapes), `lj` (generate-mode prior sampled on the GPU)."""
