SkyCoord = None
