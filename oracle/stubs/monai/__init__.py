"""Import-only stand-in (the reference imports monai at module scope in an out-of-scope discriminator)."""
