"""Host-side mirror of the reference's `lib` package, restricted to the pairwise-registration hot path."""
