"""``src.d_fine`` of the reference: model / loss / optimizer builders, matcher, criterion, dist helpers."""
