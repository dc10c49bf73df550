arcsec = deg = None
