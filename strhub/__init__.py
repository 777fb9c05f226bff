"""Import-path compatibility with baudm/parseq: `strhub.models.parseq.system.PARSeq`,
`strhub.models.utils.create_model`, `strhub.data.utils.Tokenizer` resolve to parseq_b200."""
