"""Host-side mirror of the reference's models/modules/* for the render hot path."""
