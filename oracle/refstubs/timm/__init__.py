"""Minimal stand-in for `timm`, only so that the UNMODIFIED reference (segmentation/denseclip/models.py:9,11) can be
imported in the build container to generate golden fixtures.  Test infrastructure only; never imported by the product."""
