AnchoredDirectionArrows = None
