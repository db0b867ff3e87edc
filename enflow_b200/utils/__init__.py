"""Host-side helpers of the B200 path: unit conversion (reduced Lennard-Jones units) and the small tensor utilities of
the reference's `utils.helpers`, the segment reductions among them running through the C ABI."""
