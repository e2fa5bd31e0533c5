"""Test-only stand-in for the `yacs` package (absent from this image, no network): the reference's os2d/config.py does
`from yacs.config import CfgNode`.  Only the surface the reference uses is provided."""
