Vizier = None
