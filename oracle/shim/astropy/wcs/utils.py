pixel_to_skycoord = None
