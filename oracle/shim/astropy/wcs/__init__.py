WCS = None
